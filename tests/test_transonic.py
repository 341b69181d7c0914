"""Transonic pressure equation (simple_.transonic()): pEqn = fvm::div(phid, p) - fvm::laplacian(rho rAU, p) of DATurboFoam
(reference DAResidualTurboFoam.C:148-189) and DARhoSimpleCFoam (DAResidualRhoSimpleCFoam.C:148-200): residual, transpose product,
preconditioner residual (div(pc) scheme, transonicPCOption) and adjoint solve against the oracle's restatement and its tape."""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, rel_err, synthetic_state
from tests.test_mrf import NRES_C, NS_C, mrf_spec

SCHEMES = {"Gauss upwind": (0, 1.0), "Gauss linear": (2, 1.0), "Gauss limitedLinear 1.0": (4, 1.0), "Gauss limitedLinear 0.5": (4, 0.5)}


def setup_transonic(solver, scheme, lib_path, with_mrf=False, pc_option=-1, function=None, U0=(230.0, 8.0, 0.0)):
    from dafoam_b200.pyDASolvers import pyDASolvers
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=2)
    th = cases.default_thermo(energy="sensibleEnthalpy" if solver == "DATurboFoam" else "sensibleInternalEnergy")
    bcs = cases.compressible_bcs(cases.default_bcs_naca(U0=U0))
    ns = dict(NS_C, U=float(U0[0]))
    orc = Oracle(mesh, bcs, normalizeStates=ns, normalizeResiduals=NRES_C, thermo=th, divU="linearUpwindV")
    orc.set_turbo(solver == "DATurboFoam")
    code, k = SCHEMES[scheme]
    orc.set_transonic(True, code, k, pc_option)
    W = synthetic_state(mesh, orc.geometry("C"), orc.geometry("Sf"), U0=U0, thermo=th)
    opts = dict(normalizeStates=ns, normalizeResiduals=list(NRES_C), transonicPCOption=pc_option)
    if function:
        opts["function"] = function
    kw = dict(thermo=th)
    if with_mrf:
        mrf = mrf_spec(mesh, orc.geometry("C").reshape(-1, 3), omega=40.0)
        orc.set_mrf(mesh, mrf)
        kw["mrf"] = mrf
    d = tempfile.mkdtemp(prefix="dab_transonic_")
    # DARhoSimpleCFoam is transonic whatever fvSolution says; DATurboFoam reads SIMPLE { transonic yes; }
    cases.write_case(d, mesh, bcs, div_u="bounded Gauss linearUpwindV grad(U)", transonic=(solver == "DATurboFoam"), div_phid_p=scheme, **kw)
    sol = pyDASolvers("%s -python" % solver, opts, caseDir=d, _lib_path=lib_path)
    return mesh, orc, sol, W


def test_oracle_transonic_rows_and_tape():
    """The restated rows really change (p and phi rows only), and the tape transpose agrees with central differences."""
    mesh, orc, sol, W = setup_transonic("DARhoSimpleCFoam", "Gauss limitedLinear 1.0", HOSTSIM)
    nC, nF = mesh.n_cells, mesh.n_faces
    R1 = orc.residual(W)
    orc.set_transonic(False)
    R0 = orc.residual(W)
    orc.set_transonic(True, 4, 1.0, -1)
    d = R1 - R0
    assert np.abs(d[:3 * nC]).max() == 0 and np.abs(d[4 * nC:6 * nC]).max() == 0          # U, T, nuTilda rows untouched
    assert np.linalg.norm(d[3 * nC:4 * nC]) > 1e-6 * np.linalg.norm(R0[3 * nC:4 * nC])     # p row
    assert np.linalg.norm(d[-nF:]) > 0                                                    # phi row
    # upwind vs linear differ; limitedLinear lies between them on the faces it limits
    orc.record(W)
    rng = np.random.default_rng(0)
    psi = rng.uniform(-1, 1, orc.ndof)
    g = orc.jtvec(psi, normalize=False)
    v = rng.uniform(-1, 1, orc.ndof) * np.abs(W) * 1e-2 + 1e-12
    eps = 1e-6
    fd = psi @ (orc.residual(W + eps * v) - orc.residual(W - eps * v)) / (2 * eps)
    assert abs(g @ v - fd) <= 2e-6 * abs(fd), (g @ v, fd)


def check_engine(lib_path, tol=1e-10):
    for solver, scheme, with_mrf in (("DARhoSimpleCFoam", "Gauss upwind", False), ("DARhoSimpleCFoam", "Gauss linear", False),
                                     ("DARhoSimpleCFoam", "Gauss limitedLinear 1.0", False), ("DATurboFoam", "Gauss limitedLinear 0.5", True),
                                     ("DATurboFoam", "Gauss upwind", True)):
        mesh, orc, sol, W = setup_transonic(solver, scheme, lib_path, with_mrf=with_mrf)
        sol.updateOFFields(W)
        R = np.zeros(orc.ndof)
        sol.getResiduals(R)
        assert rel_err(R, orc.residual(W)) < tol, (solver, scheme, rel_err(R, orc.residual(W)))
        orc.record(W)
        rng = np.random.default_rng(7)
        for _ in range(2):
            psi = rng.uniform(-1, 1, orc.ndof)
            y = np.zeros(orc.ndof)
            sol.calcdRdWTPsiAD(psi, y)
            e = rel_err(y, orc.jtvec(psi))
            assert e < tol, (solver, scheme, e)
        # the limiter is active on a fair share of the faces (neither everywhere clipped nor nowhere)
        # preconditioner residual: div(pc) = upwind for div(phid,p)
        Rpc = np.zeros(orc.ndof)
        sol.getResiduals(Rpc, isPC=1)
        assert rel_err(Rpc, orc.residual(W, isPC=1)) < tol, (solver, scheme)


def test_transonic_residual_and_transpose_product_host_build():
    check_engine(HOSTSIM)


@pytest.mark.gpu
def test_transonic_residual_and_transpose_product_cuda():
    check_engine(None, tol=1e-9)


def check_pc_options_and_adjoint(lib_path, tol=1e-10):
    # transonicPCOption 1 / 2 only change the preconditioner residual
    for solver, opt in (("DARhoSimpleCFoam", 1), ("DATurboFoam", 1), ("DATurboFoam", 2)):
        mesh, orc, sol, W = setup_transonic(solver, "Gauss limitedLinear 1.0", lib_path, with_mrf=(solver == "DATurboFoam"), pc_option=opt)
        sol.updateOFFields(W)
        R, Rpc = np.zeros(orc.ndof), np.zeros(orc.ndof)
        sol.getResiduals(R)
        sol.getResiduals(Rpc, isPC=1)
        assert rel_err(R, orc.residual(W)) < tol
        assert rel_err(Rpc, orc.residual(W, isPC=1)) < tol, (solver, opt, rel_err(Rpc, orc.residual(W, isPC=1)))
    # adjoint solve of a force with the transonic rows: GMRES on the engine's product and preconditioner vs the oracle's J^T
    from dafoam_b200.pyDASolvers import KSP, Mat
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection", "direction": [1.0, 0.0, 0.0],
                 "scale": 1e-3}}
    mesh, orc, sol, W = setup_transonic("DARhoSimpleCFoam", "Gauss limitedLinear 1.0", lib_path, pc_option=1, function=fn)
    sol.updateDAOption(dict(adjEqnOption=dict(gmresRelTol=1e-10, gmresMaxIters=600, gmresRestart=600, pcConLevel=3)))
    sol.updateOFFields(W)
    n = orc.ndof
    dFdW = np.zeros(n)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    psi = np.zeros(n)
    assert sol.solveLinearEqn(ksp, dFdW, psi) == 0
    orc.record(W)
    assert rel_err(orc.jtvec(psi), dFdW) < 1e-7
    with pytest.raises(Exception, match="transonic pressure corrector"):
        sol.solvePrimal()


def test_transonic_pc_options_and_adjoint_host_build():
    check_pc_options_and_adjoint(HOSTSIM)


@pytest.mark.gpu
def test_transonic_pc_options_and_adjoint_cuda():
    check_pc_options_and_adjoint(None, tol=1e-9)
