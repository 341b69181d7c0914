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


# ---- tightening of the (unpinned) oracle: exact tangents and closed-form operator identities --------------------------------
HOSTSIM_LIB = __import__("tests.common", fromlist=["HOSTSIM"]).HOSTSIM


@pytest.mark.parametrize("kind,turb,divU", [("naca", True, "linearUpwindV"), ("channel", True, "linearUpwind"), ("channel", False, "upwind"),
                                            ("nacawf", True, "linearUpwindV"), ("prism", True, "linearUpwind")])
def test_tape_equals_dual_number_tangent(kind, turb, divU):
    """psi^T (J v) from the forward dual-number instantiation of the residual equals v^T (J^T psi) from the reverse tape to
    rounding -- two independent differentiations of the same restatement (central differences only reach 1e-7)."""
    mesh, bcs, orc, sol, W, _ = setup(kind, turb, divU=divU, nk=1, lib_path=HOSTSIM_LIB)
    rng = np.random.default_rng(17)
    for isPC in (0, 1):
        orc.record(W, isPC)
        psi = rng.uniform(-1, 1, orc.ndof)
        y = orc.jtvec(psi, normalize=False)
        v = rng.uniform(-1, 1, orc.ndof)
        Jv = orc.jvec(W, v, isPC)
        lhs, rhs = psi @ Jv, v @ y
        assert abs(lhs - rhs) <= 1e-11 * max(abs(rhs), np.linalg.norm(Jv) * np.linalg.norm(psi) * 1e-3), (isPC, lhs, rhs)


def test_tape_equals_dual_number_tangent_compressible():
    from dafoam_b200 import cases
    from oracle.pyoracle import Oracle, synthetic_state
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=1)
    ns = dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0)
    th = cases.default_thermo(energy="sensibleEnthalpy", transport="sutherland")
    bcs = cases.compressible_bcs(cases.default_bcs_naca(U0=(50.0, 2.0, 0.0)))
    orc = Oracle(mesh, bcs, normalizeStates=ns, normalizeResiduals=("URes", "pRes", "TRes", "nuTildaRes", "phiRes"), thermo=th)
    W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=(50.0, 2.0, 0.0), thermo=th)
    orc.record(W)
    rng = np.random.default_rng(2)
    psi = rng.uniform(-1, 1, orc.ndof)
    v = rng.uniform(-1, 1, orc.ndof) * (np.abs(W) * 1e-2 + 1e-9)
    lhs, rhs = psi @ orc.jvec(W, v), v @ orc.jtvec(psi, normalize=False)
    assert abs(lhs - rhs) <= 1e-11 * abs(rhs)


def _uniform_channel(turbulent, nx=9, ny=8):
    """Orthogonal, uniform channel (no contraction, no shear) + its oracle with per-volume residuals; `inner` = cells at least two
    cells away from every x/y boundary (their stencils, and those of their neighbours, see no boundary face but the symmetry planes)."""
    from dafoam_b200 import cases
    from oracle.pyoracle import Oracle
    mesh = cases.channel(nx=nx, ny=ny, nz=1, contraction=0.0, skew=0.0)
    bcs = cases.default_bcs_channel(turbulent=turbulent)
    nres = ("URes", "pRes", "nuTildaRes", "phiRes")
    orc = Oracle(mesh, bcs, normalizeStates=dict(U=1.0, p=1.0, nuTilda=1.0, phi=1.0), divU="linearUpwind", normalizeResiduals=nres)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    cid = (i + nx * j).ravel()
    inner = cid[((i >= 2) & (i < nx - 2) & (j >= 2) & (j < ny - 2)).ravel()]
    return mesh, orc, inner


def test_gauss_gradient_and_laplacian_are_exact_for_a_linear_pressure():
    """fvc::grad(p) (Gauss linear) of a linear field is the constant gradient, and laplacian(rAU, p) of it vanishes, on an orthogonal
    uniform mesh: with U = 0 and phi = 0, URes = grad(p) and pRes = 0 in the interior (Appendix A.4 of SURVEY.md)."""
    mesh, orc, inner = _uniform_channel(False)
    nC = mesh.n_cells
    C = orc.geometry("C").reshape(nC, 3)
    Cx, Cy = C[:, 0], C[:, 1]
    a, b = 3.0, -2.0
    W = np.zeros(orc.ndof)
    W[3 * nC:4 * nC] = a * Cx + b * Cy + 0.7
    R = orc.residual(W)
    URes = R[:3 * nC].reshape(nC, 3)
    assert np.allclose(URes[inner, 0], a, rtol=0, atol=1e-11)
    assert np.allclose(URes[inner, 1], b, rtol=0, atol=1e-11)
    assert np.allclose(URes[inner, 2], 0.0, rtol=0, atol=1e-11)
    # the inlet's fixedValue U (10 m/s against U = 0 in the cells) reaches HbyA of the second cell column through the explicit dev2
    # term: the pressure rows are checked one more column away from the inlet
    pRes = R[3 * nC:4 * nC]
    inner3 = inner[(inner % 9) >= 3]
    assert inner3.size >= 12 and np.abs(pRes[inner3]).max() <= 1e-12 * max(1.0, np.abs(pRes).max())


def test_bounded_convection_of_a_constant_velocity_vanishes():
    """`bounded Gauss`: div(phi,U) - Sp(div(phi)) U of a constant U is zero for ANY face flux (conservative or not); diffusion and the
    dev2 term of a constant vanish too, so URes = 0 in the interior with p = 0."""
    mesh, orc, inner = _uniform_channel(False)
    nC = mesh.n_cells
    W = np.zeros(orc.ndof)
    W[:3 * nC] = np.tile([4.0, -1.5, 0.0], nC)
    W[4 * nC:] = np.random.default_rng(5).uniform(-1e-3, 1e-3, mesh.n_faces)  # arbitrary, non-conservative flux
    URes = orc.residual(W)[:3 * nC].reshape(nC, 3)
    assert np.abs(URes[inner]).max() <= 1e-12 * max(1.0, np.abs(URes).max())


def test_sa_source_terms_at_a_hand_computed_state():
    """Uniform nuTilda in a plane shear flow u = gamma*y with phi = 0: convection, diffusion and the Cb2 term vanish and
    nuTildaRes = -Cb1 Stilda nuTilda + Cw1 fw (nuTilda/y)^2 per unit volume, with chi, fv1, fv2, Stilda, r, g, fw evaluated by hand
    from DASpalartAllmaras.C:124-178, 452-485."""
    mesh, orc, inner = _uniform_channel(True)
    nC = mesh.n_cells
    C = orc.geometry("C").reshape(nC, 3)
    yw = orc.geometry("yWall")
    gamma, nt, nu = 40.0, 3.0e-4, 1.5e-5
    W = np.zeros(orc.ndof)
    W[0:3 * nC:3] = gamma * C[:, 1]
    W[4 * nC:5 * nC] = nt
    R = orc.residual(W)[4 * nC:5 * nC]
    sigma, kappa, Cb1, Cb2, Cw2, Cw3, Cv1, Cs = 0.66666, 0.41, 0.1355, 0.622, 0.3, 2.0, 7.1, 0.3
    Cw1 = Cb1 / kappa**2 + (1.0 + Cb2) / sigma
    chi = nt / nu
    fv1 = chi**3 / (chi**3 + Cv1**3)
    fv2 = 1.0 - chi / (1.0 + chi * fv1)
    y = yw[inner]
    Omega = gamma
    St = np.maximum(Omega + fv2 * nt / (kappa * y)**2, Cs * Omega)
    r = np.minimum(nt / (np.maximum(St, 1e-15) * (kappa * y)**2), 10.0)
    g = r + Cw2 * (r**6 - r)
    fw = g * ((1.0 + Cw3**6) / (g**6 + Cw3**6))**(1.0 / 6.0)
    expect = -Cb1 * St * nt + Cw1 * fw * (nt / y)**2
    assert np.allclose(R[inner], expect, rtol=1e-11, atol=0)
