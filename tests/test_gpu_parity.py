"""-m gpu: the parity tests proper, through the C ABI of the CUDA library (libdab200.so) on a B200."""
import numpy as np
import pytest

from tests.common import check_functions, check_parity, rel_err, setup

pytestmark = pytest.mark.gpu


def test_forward_and_reverse_parity_cuda():
    worst = check_parity(None, tol=1e-10)
    print("worst relative difference vs oracle: %.3e" % worst)
    assert worst < 1e-10


def test_force_function_and_dfdw_cuda():
    dirv = [float(np.cos(0.05)), float(np.sin(0.05)), 0.0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": dirv, "scale": 0.02}}
    mesh, bcs, orc, sol, W, _ = setup("naca", True, extra_options=dict(function=fn))
    sol.updateOFFields(W)
    F, Fo = sol.calcFunction("CD"), orc.force(W, 0, dirv, 0.02)
    assert abs(F - Fo) <= 1e-12 * abs(Fo)
    prod = np.zeros(orc.ndof)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), prod)
    assert rel_err(prod, orc.dforce_dw(W, 0, dirv, 0.02)) < 1e-12


def test_force_moment_and_direction_modes_cuda():
    assert check_functions(None)


def test_full_size_linearity_and_determinism_cuda():
    # BASELINE config sizes are beyond the oracle: use size-independent properties of the linear operator
    from dafoam_b200 import cases
    from dafoam_b200.pyDASolvers import pyDASolvers
    import tempfile
    mesh = cases.naca0012_ogrid(ni=1400, nj=700, nk=1)
    d = tempfile.mkdtemp(prefix="dab_big_")
    cases.write_case(d, mesh, cases.default_bcs_naca(), binary=True)
    sol = pyDASolvers("DASimpleFoam -python", dict(normalizeStates=dict(U=10.0, p=50.0, nuTilda=1e-3, phi=1.0)), caseDir=d)
    n = sol.getNLocalAdjointStates()
    assert sol.getNLocalCells() == 980000
    rng = np.random.default_rng(1)
    W = np.zeros(n)
    sol.getOFFields(W)
    W *= 1.0 + 0.01 * rng.uniform(-1, 1, n)
    sol.updateOFFields(W)
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    ya, yb, yab, ya2 = (np.zeros(n) for _ in range(4))
    sol.calcdRdWTPsiAD(a, ya)
    sol.calcdRdWTPsiAD(b, yb)
    sol.calcdRdWTPsiAD(2.0 * a - 3.0 * b, yab)
    sol.calcdRdWTPsiAD(a, ya2)
    assert np.array_equal(ya, ya2)  # gather kernels: bitwise reproducible
    assert rel_err(yab, 2.0 * ya - 3.0 * yb) < 1e-12
    assert np.isfinite(ya).all() and np.linalg.norm(ya) > 0
