"""solvePrimal (device SIMPLE, reference DASimpleFoam::solvePrimal): the fixed point must be the root of the same R(W)
the adjoint differentiates.  Checked on the host build (CPU suite) against the oracle: (i) the oracle's residual of the
SIMPLE state vanishes, (ii) the state equals the oracle's own Newton-converged state, and (iii) the whole reference
workflow solvePrimal -> solveAdjoint -> total derivative agrees with finite differences over re-converged primals."""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import KSP, Mat, pyDASolvers
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, NORM_STATES
from tests.test_converged_primal import newton

FN = {"CD": {"type": "force", "source": "patchToFace", "patches": ["walls"], "directionMode": "fixedDirection",
             "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
INP = {"patchV": {"type": "patchVelocity", "patches": ["inlet"], "flowAxis": "x", "normalAxis": "y"}}


def make(lib_path, tol=1e-12, **dicts):
    mesh, bcs = cases.channel(nx=14, ny=8, nz=1), cases.default_bcs_channel()
    d = tempfile.mkdtemp(prefix="dab_primal_")
    cases.write_case(d, mesh, bcs, **dicts)
    opts = dict(normalizeStates=NORM_STATES, function=FN, inputInfo=INP, primalMinResTol=tol, primalMaxIters=2000,
                adjEqnOption=dict(gmresRelTol=1e-12, gmresMaxIters=400, gmresRestart=400, pcConLevel=3))
    sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=lib_path)
    return mesh, bcs, sol


def run_fixed_point(lib_path):
    mesh, bcs, sol = make(lib_path)
    n = sol.getNLocalAdjointStates()
    W0 = np.zeros(n)
    sol.getOFFields(W0)
    assert sol.solvePrimal() == 0
    st = sol.primalStats
    assert st.converged == 1 and st.max_residual < 1e-12 and 10 < st.iterations < 2000
    W = np.zeros(n)
    sol.getOFFields(W)
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES)
    r0, r1 = np.linalg.norm(orc.residual(W0)), np.linalg.norm(orc.residual(W))
    assert r1 < 1e-9 * r0, (r0, r1)
    Wn = newton(orc, W0.copy(), tol=1e-9)
    nC = mesh.n_cells
    for name, a, b in (("U", 0, 3 * nC), ("p", 3 * nC, 4 * nC), ("nuTilda", 4 * nC, 5 * nC), ("phi", 5 * nC, n)):
        err = np.linalg.norm(W[a:b] - Wn[a:b]) / np.linalg.norm(Wn[a:b])
        assert err < 1e-8, (name, err)
    # a second call starts from the converged state: only the round-off tail is left
    assert sol.solvePrimal() == 0 and sol.primalStats.iterations <= 40


def run_simplec(lib_path):
    """fvSolution SIMPLE { consistent yes; } (SIMPLEC, reference pEqnSimple.H:27-33): the same fixed point, reached in
    fewer iterations without pressure under-relaxation."""
    mesh, bcs, sol = make(lib_path)
    assert sol.solvePrimal() == 0
    itSimple = sol.primalStats.iterations
    n = sol.getNLocalAdjointStates()
    W = np.zeros(n)
    sol.getOFFields(W)
    mesh, bcs, solc = make(lib_path, consistent=True, relax_u=0.7, relax_p=1.0)
    assert solc.solvePrimal() == 0
    st = solc.primalStats
    assert st.converged == 1 and st.max_residual < 1e-12
    Wc = np.zeros(n)
    solc.getOFFields(Wc)
    nC = mesh.n_cells
    for name, a, b in (("U", 0, 3 * nC), ("p", 3 * nC, 4 * nC), ("nuTilda", 4 * nC, 5 * nC), ("phi", 5 * nC, n)):
        err = np.linalg.norm(W[a:b] - Wc[a:b]) / np.linalg.norm(W[a:b])
        assert err < 1e-8, (name, err)
    assert st.iterations < 0.8 * itSimple, (st.iterations, itSimple)


def run_workflow(lib_path):
    """solvePrimal -> dF/dW -> solveAdjoint -> total derivative w.r.t. (|U|, aoa), against FD of re-converged primals."""
    mesh, bcs, sol = make(lib_path)
    n = sol.getNLocalAdjointStates()
    x0 = np.array([10.0, 2.0])

    def F_at(x):
        sol.setSolverInput("patchV", "patchVelocity", 2, x)
        assert sol.solvePrimal() == 0
        return sol.calcFunction("CD")

    F0 = F_at(x0)
    W = np.zeros(n)
    sol.getOFFields(W)
    one = np.array([1.0])
    dFdW, dFdx, psi, prod = np.zeros(n), np.zeros(2), np.zeros(n), np.zeros(2)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", one, dFdW)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x0, "CD", "function", one, dFdx)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    assert sol.solveLinearEqn(ksp, dFdW, psi) == 0
    sol.calcJacTVecProduct("patchV", "patchVelocity", x0, "R", "residual", psi, prod)
    total = dFdx - prod
    fd = np.zeros(2)
    for k, h in enumerate((1e-3, 1e-2)):
        xp, xm = x0.copy(), x0.copy()
        xp[k] += h
        xm[k] -= h
        fd[k] = (F_at(xp) - F_at(xm)) / (2 * h)
    assert np.isfinite(F0)
    assert np.allclose(total, fd, rtol=2e-5, atol=0.0), (total, fd)


def run_shape_workflow(lib_path):
    """Shape derivative, the reference's main use (DAFoamSolver.linearize/apply_linear + DAInputVolCoord): the total
    dF/d(alpha) of a mesh deformation x = x0 + alpha*v, as dF/dx.v - psi^T dR/dx.v, against FD over deformed meshes with
    re-converged primals."""
    mesh, bcs, sol = make(lib_path)
    n, nP3 = sol.getNLocalAdjointStates(), 3 * sol.getNLocalPoints()
    x0 = np.zeros(nP3)
    sol.getOFMeshPoints(x0)
    X = x0.reshape(-1, 3)
    ymin, ymax, xmin, xmax = X[:, 1].min(), X[:, 1].max(), X[:, 0].min(), X[:, 0].max()
    s = (X[:, 0] - xmin) / (xmax - xmin)
    t = (X[:, 1] - ymin) / (ymax - ymin)
    v = np.zeros_like(X)
    v[:, 1] = 0.05 * (ymax - ymin) * np.sin(np.pi * s) ** 2 * (1.0 - t)  # a bump on the lower wall, fading to the upper one
    v = v.ravel()

    def F_at(alpha):
        sol.updateOFMesh(x0 + alpha * v)
        assert sol.solvePrimal() == 0
        return sol.calcFunction("CD")

    F0 = F_at(0.0)
    W = np.zeros(n)
    sol.getOFFields(W)
    one = np.array([1.0])
    dFdW, psi, dFdx, dRdxTpsi = np.zeros(n), np.zeros(n), np.zeros(nP3), np.zeros(nP3)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", one, dFdW)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    assert sol.solveLinearEqn(ksp, dFdW, psi) == 0
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", x0, "CD", "function", one, dFdx)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", x0, "R", "residual", psi, dRdxTpsi)
    total = (dFdx - dRdxTpsi) @ v
    h = 1e-3
    fd = (F_at(h) - F_at(-h)) / (2 * h)
    sol.updateOFMesh(x0)
    assert np.isfinite(F0) and abs(fd) > 0
    assert abs(total - fd) <= 2e-5 * abs(fd), (total, fd)


def run_pydafoam_api(lib_path):
    """The user-facing PYDAFOAM class (reference dafoam/pyDAFoam.py): primal, functions, adjoint, total derivatives."""
    from dafoam_b200.pyDAFoam import PYDAFOAM
    mesh, bcs = cases.channel(nx=14, ny=8, nz=1), cases.default_bcs_channel()
    d = tempfile.mkdtemp(prefix="dab_pydafoam_")
    cases.write_case(d, mesh, bcs)
    inp = dict(INP, aero_vol_coords={"type": "volCoord", "components": ["solver", "function"]})
    opts = dict(solverName="DASimpleFoam", normalizeStates=NORM_STATES, function=FN, inputInfo=inp, primalMinResTol=1e-12, primalMaxIters=2000,
                adjEqnOption=dict(gmresRelTol=1e-12, gmresMaxIters=400, gmresRestart=400, pcConLevel=3))
    DASolver = PYDAFOAM(options=opts, comm=None, caseDir=d, _lib_path=lib_path)
    x = np.array([10.0, 2.0])
    DASolver.set_solver_input({"patchV": x})
    DASolver()
    assert DASolver.primalFail == 0
    funcs = {}
    DASolver.evalFunctions(funcs)
    assert set(funcs) == {"CD"} and np.isfinite(funcs["CD"])
    assert np.linalg.norm(DASolver.getResiduals()) < 1e-6
    total = DASolver.calcTotalDeriv("CD", "patchV", x)
    assert DASolver.adjointFail == 0
    h = 1e-3
    fd = np.zeros(2)
    for k in range(2):
        f2 = []
        for sgn in (1.0, -1.0):
            xp = x.copy()
            xp[k] += sgn * h
            DASolver.set_solver_input({"patchV": xp})
            DASolver()
            fk = {}
            DASolver.evalFunctions(fk)
            f2.append(fk["CD"])
        fd[k] = (f2[0] - f2[1]) / (2 * h)
    assert np.allclose(total, fd, rtol=5e-5), (total, fd)
    DASolver.set_solver_input({"patchV": x})
    DASolver()
    dFdxv = DASolver.calcTotalDeriv("CD", "aero_vol_coords")
    assert dFdxv.shape == (3 * DASolver.getNLocalPoints(),) and np.linalg.norm(dFdxv) > 0
    # adjPCLag (reference mphys_dafoam.py:511-530): the preconditioner of an earlier design keeps being used; the adjoint still
    # converges to the same derivative, the assembly count only moves every adjPCLag designs
    n0 = DASolver._ksp.stats.pc_assemblies
    DASolver.setOption("adjPCLag", 2)
    DASolver.updateDAOption()
    counts, totals = [], []
    c0 = DASolver.solution_counter
    for k in range(3):
        xk = x + np.array([0.2 * (k + 1), 0.1 * (k + 1)])
        DASolver.set_solver_input({"patchV": xk})
        DASolver()
        totals.append(DASolver.calcTotalDeriv("CD", "patchV", xk))
        assert DASolver.adjointFail == 0
        counts.append(DASolver._ksp.stats.pc_assemblies - n0)
    # re-assembled in the derivative iterations with (solution_counter - 1) % adjPCLag == 0, kept in between
    expected = [int(v) for v in np.cumsum([1 if (c0 + k) % 2 == 0 else 0 for k in range(3)])]
    assert counts == expected and counts[-1] in (1, 2), (counts, expected)
    DASolver.setOption("adjPCLag", 1)
    DASolver.updateDAOption()
    DASolver.set_solver_input({"patchV": x + np.array([0.4, 0.2])})
    DASolver()
    fresh = DASolver.calcTotalDeriv("CD", "patchV", x + np.array([0.4, 0.2]))
    assert np.allclose(totals[1], fresh, rtol=1e-7), (totals[1], fresh)


def test_pydafoam_class_host_build():
    run_pydafoam_api(HOSTSIM)


@pytest.mark.gpu
def test_pydafoam_class_cuda():
    run_pydafoam_api(None)


def test_simple_fixed_point_is_the_root_of_the_residual_host_build():
    run_fixed_point(HOSTSIM)


def test_simplec_reaches_the_same_fixed_point_host_build():
    run_simplec(HOSTSIM)


@pytest.mark.gpu
def test_simplec_reaches_the_same_fixed_point_cuda():
    run_simplec(None)


def test_primal_adjoint_workflow_matches_fd_host_build():
    run_workflow(HOSTSIM)


def test_shape_derivative_workflow_matches_fd_host_build():
    run_shape_workflow(HOSTSIM)


@pytest.mark.gpu
def test_shape_derivative_workflow_matches_fd_cuda():
    run_shape_workflow(None)


@pytest.mark.gpu
def test_simple_fixed_point_is_the_root_of_the_residual_cuda():
    run_fixed_point(None)


@pytest.mark.gpu
def test_primal_adjoint_workflow_matches_fd_cuda():
    run_workflow(None)
