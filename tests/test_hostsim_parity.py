"""Kernel functors (compiled for the host: tests/hostsim, TEST-ONLY) vs the oracle.  This is how the
hand-derived reverse sweep is checked on a machine without a GPU; the -m gpu tests repeat the same
checks through the CUDA library."""
import numpy as np

from tests.common import HOSTSIM, check_functions, check_parity, setup, rel_err


def test_forward_and_reverse_parity_host_build():
    worst = check_parity(HOSTSIM, tol=1e-10)
    assert worst < 1e-10


def test_force_function_and_dfdw_host_build():
    dirv = [float(np.cos(0.05)), float(np.sin(0.05)), 0.0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": dirv, "scale": 0.02}}
    mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=HOSTSIM, extra_options=dict(function=fn))
    sol.updateOFFields(W)
    F, Fo = sol.calcFunction("CD"), orc.force(W, 0, dirv, 0.02)
    assert abs(F - Fo) <= 1e-12 * abs(Fo)
    prod = np.zeros(orc.ndof)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), prod)
    assert rel_err(prod, orc.dforce_dw(W, 0, dirv, 0.02)) < 1e-12


def test_force_moment_and_direction_modes_host_build():
    assert check_functions(HOSTSIM)


def test_dot_product_identity_host_build():
    # <J^T psi, v> == <psi, J v> with J v from central differences of the engine's own R(W)
    mesh, bcs, orc, sol, W, _ = setup("channel", True, lib_path=HOSTSIM, nk=1,
                                      extra_options=dict(normalizeStates=dict(U=1.0, p=1.0, nuTilda=1.0, phi=1.0)))
    sol.updateOFFields(W)
    rng = np.random.default_rng(5)
    n = orc.ndof
    psi = rng.uniform(-1, 1, n)
    y = np.zeros(n)
    sol.calcdRdWTPsiAD(psi, y)
    magSf = orc.geometry("magSf")
    y[-mesh.n_faces:] /= magSf  # undo the phi scaling (normalizeStates phi * magSf)
    v = rng.uniform(-1, 1, n) * 1e-3
    eps = 1e-4
    Rp, Rm = np.zeros(n), np.zeros(n)
    sol.updateOFFields(W + eps * v)
    sol.getResiduals(Rp)
    sol.updateOFFields(W - eps * v)
    sol.getResiduals(Rm)
    Jv = (Rp - Rm) / (2 * eps)
    assert abs(psi @ Jv - v @ y) <= 1e-6 * abs(v @ y)
