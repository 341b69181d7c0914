"""The reference validates its adjoint by comparing total derivatives with a second, independent differentiation of the
CONVERGED primal (forward-mode AD per design variable: tests/testFuncs.py:17-52, tests/refs/DAFoam_Test_DASimpleFoamForwardRef.txt).
Same check here: a primal state converged to |R| < 1e-10 (Newton on the oracle's residual -- test infrastructure; the
engine has no primal solver yet), then
    adjoint total  dF/dx = dF/dx|_W - psi^T dR/dx   (engine: GMRES on the GPU/host build)
    vs central finite differences of F(W*(x)) over re-converged primals."""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import KSP, Mat, pyDASolvers
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, NORM_STATES


def newton(orc, W, tol=1e-10, maxit=40):
    """Damped Newton on R(W) = 0 with the oracle's exact Jacobian (rows from tape transposes)."""
    n = orc.ndof
    for _ in range(maxit):
        R = orc.residual(W)
        nr = np.linalg.norm(R)
        if nr < tol:
            return W
        orc.record(W)
        J = np.zeros((n, n))
        e = np.zeros(n)
        for i in range(n):
            e[:] = 0.0
            e[i] = 1.0
            J[i] = orc.jtvec(e, normalize=False)
        d = np.linalg.solve(J, -R)
        lam = 1.0
        while lam > 1e-4:
            Rn = orc.residual(W + lam * d)
            if np.all(np.isfinite(Rn)) and np.linalg.norm(Rn) < (1.0 - 1e-4 * lam) * nr:
                break
            lam *= 0.5
        W = W + lam * d
    raise AssertionError("primal Newton iteration did not converge")


def run(lib_path):
    mesh, bcs = cases.channel(nx=14, ny=8, nz=1), cases.default_bcs_channel()
    names = [p["name"] for p in mesh.patches]
    ip, fp = names.index("inlet"), names.index("walls")
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES)
    n, nC = orc.ndof, mesh.n_cells
    x0 = np.array([10.0, 2.0])  # |U|, angle of attack [deg]

    def set_x(x):
        a = np.deg2rad(x[1])
        orc.set_bc_value("U", ip, [x[0] * np.cos(a), x[0] * np.sin(a), 0.0])

    def F_of(W):
        return orc.force(W, fp, [1.0, 0.0, 0.0], 1.0)

    Sf = orc.geometry("Sf").reshape(-1, 3)
    W = np.concatenate([np.tile([10.0, 0.0, 0.0], nC), np.zeros(nC), np.full(nC, 4.5e-5), 10.0 * Sf[:, 0]])
    for pch in mesh.patches:
        if pch["type"] in ("symmetry", "wall"):
            W[5 * nC + pch["start"]:5 * nC + pch["start"] + pch["size"]] = 0.0
    set_x(x0)
    Wc = newton(orc, W)
    assert np.linalg.norm(orc.residual(Wc)) < 1e-10
    # central differences over re-converged primals
    fd = np.zeros(2)
    for k, h in enumerate((1e-4, 1e-3)):
        xp, xm = x0.copy(), x0.copy()
        xp[k] += h
        xm[k] -= h
        set_x(xp)
        Fp = F_of(newton(orc, Wc))
        set_x(xm)
        Fm = F_of(newton(orc, Wc))
        fd[k] = (Fp - Fm) / (2 * h)
    set_x(x0)
    # engine adjoint at the converged state
    d = tempfile.mkdtemp(prefix="dab_conv_")
    cases.write_case(d, mesh, bcs)
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["walls"], "directionMode": "fixedDirection",
                 "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
    inp = {"patchV": {"type": "patchVelocity", "patches": ["inlet"], "flowAxis": "x", "normalAxis": "y"}}
    opts = dict(normalizeStates=NORM_STATES, function=fn, inputInfo=inp,
                adjEqnOption=dict(gmresRelTol=1e-12, gmresMaxIters=900, gmresRestart=300))
    sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=lib_path)
    sol.updateOFFields(Wc)
    dFdx = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x0, "CD", "function", np.array([1.0]), dFdx)
    R = np.zeros(n)
    sol.getResiduals(R)
    assert np.linalg.norm(R) < 1e-9  # the engine agrees that the state is converged
    assert abs(sol.calcFunction("CD") - F_of(Wc)) <= 1e-12 * abs(F_of(Wc))
    b = np.zeros(n)
    sol.calcJacTVecProduct("states", "stateVar", Wc, "CD", "function", np.array([1.0]), b)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    psi = np.zeros(n)
    assert sol.solveLinearEqn(ksp, b, psi) == 0
    prod = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x0, "R", "residual", psi, prod)
    total = dFdx - prod
    return total, fd


def check(lib_path):
    total, fd = run(lib_path)
    assert np.all(np.abs(fd) > 0)
    # adjoint vs finite differences of the converged primal: limited by the FD error (step 1e-4 on |U|: ~1e-7;
    # step 1e-3 deg on the angle whose derivative is 4000x smaller: ~5e-6)
    assert abs(total[0] - fd[0]) <= 1e-6 * abs(fd[0]), (total, fd)
    assert abs(total[1] - fd[1]) <= 5e-5 * abs(fd[1]), (total, fd)


def test_adjoint_total_matches_fd_of_converged_primal_host_build():
    check(HOSTSIM)


@pytest.mark.gpu
def test_adjoint_total_matches_fd_of_converged_primal_cuda():
    check(None)
