"""The oracle against itself (parity with the reference is unpinned: SURVEY.md section 8c):
tape-AD transpose products vs central finite differences of R(W), and the dot-product identity."""
import numpy as np
import pytest

from tests.common import setup, segments


@pytest.mark.parametrize("kind,turb", [("naca", True), ("channel", True), ("channel", False)])
def test_tape_matches_finite_differences(kind, turb):
    mesh, bcs, orc, sol, W, _ = setup(kind, turb, nk=1, lib_path=__import__("tests.common", fromlist=["HOSTSIM"]).HOSTSIM)
    orc.record(W)
    rng = np.random.default_rng(7)
    psi = rng.uniform(-1, 1, orc.ndof)
    y = orc.jtvec(psi, normalize=False)
    nC = mesh.n_cells
    scale = np.ones(orc.ndof)
    for name, a, b in segments(mesh, turb, orc.ndof):
        scale[a:b] = {"U": 10.0, "p": 50.0, "nuTilda": 1e-4, "phi": 1e-3}[name]
    v = rng.uniform(-1, 1, orc.ndof) * scale
    eps = 1e-6
    Jv = (orc.residual(W + eps * v) - orc.residual(W - eps * v)) / (2 * eps)
    lhs, rhs = psi @ Jv, v @ y
    assert abs(lhs - rhs) <= 1e-7 * abs(rhs)


def test_residual_is_linear_in_p_for_fixed_rest():
    # pRes/URes/phiRes are affine in p: R(W + a dp) - R(W) = a (R(W + dp) - R(W))
    mesh, bcs, orc, sol, W, _ = setup("channel", True, nk=1, lib_path=__import__("tests.common", fromlist=["HOSTSIM"]).HOSTSIM)
    nC = mesh.n_cells
    dp = np.zeros(orc.ndof)
    dp[3 * nC:4 * nC] = np.random.default_rng(3).uniform(-1, 1, nC)
    R0 = orc.residual(W)
    d1 = orc.residual(W + dp) - R0
    d2 = orc.residual(W + 2.5 * dp) - R0
    assert np.linalg.norm(d2 - 2.5 * d1) <= 1e-10 * np.linalg.norm(d2)


def test_force_derivative_matches_fd():
    mesh, bcs, orc, sol, W, _ = setup("naca", True, nk=1, lib_path=__import__("tests.common", fromlist=["HOSTSIM"]).HOSTSIM)
    d = [1.0, 0.0, 0.0]
    g = orc.dforce_dw(W, 0, d, 1.0, normalize=False)
    rng = np.random.default_rng(11)
    v = np.zeros(orc.ndof)
    nC = mesh.n_cells
    v[:4 * nC] = rng.uniform(-1, 1, 4 * nC)
    eps = 1e-6
    fd = (orc.force(W + eps * v, 0, d) - orc.force(W - eps * v, 0, d)) / (2 * eps)
    assert abs(fd - g @ v) <= 1e-7 * abs(fd)


def test_compressible_oracle_tape_matches_fd():
    """DARhoSimpleFoam restatement (oracle.cpp residualComp): state [U|p|T|nuTilda|phi], both energy forms, const and
    sutherland transport; the tape's transpose product against central differences of the residual."""
    from dafoam_b200 import cases
    from oracle.pyoracle import Oracle, synthetic_state
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=1)
    ns = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
    nres = ("URes", "pRes", "TRes", "nuTildaRes", "phiRes")
    for energy, transport, ras in (("sensibleInternalEnergy", "const", "SpalartAllmaras"),
                                   ("sensibleEnthalpy", "sutherland", "SpalartAllmarasFv3")):
        th = cases.default_thermo(energy=energy, transport=transport)
        bcs = cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0)))
        orc = Oracle(mesh, bcs, normalizeStates=ns, normalizeResiduals=nres, thermo=th, rasModel=ras)
        nC = mesh.n_cells
        assert orc.ndof == 6 * nC + mesh.n_faces
        W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
        R = orc.residual(W)
        assert np.isfinite(R).all() and all(np.linalg.norm(R[a * nC:(a + 1) * nC]) > 0 for a in (3, 4, 5))
        orc.record(W)
        rng = np.random.default_rng(0)
        psi = rng.uniform(-1, 1, orc.ndof)
        g = orc.jtvec(psi, normalize=False)
        v = rng.uniform(-1, 1, orc.ndof) * np.abs(W) * 1e-1 + 1e-12
        eps = 1e-6
        fd = psi @ (orc.residual(W + eps * v) - orc.residual(W - eps * v)) / (2 * eps)
        assert abs(g @ v - fd) <= 1e-7 * abs(fd), (energy, g @ v, fd)
