"""Adjoint total derivative dF/d(|U|, angle of attack) = dF/dx - psi^T dR/dx (reference: OpenMDAO assembles it from
calcJacTVecProduct pieces, dafoam/mphys/mphys_dafoam.py:375-431, 746-801) -- engine vs oracle, 1e-6 relative
(the tolerance BASELINE.json's north_star states)."""
import tempfile

import numpy as np
import pytest

from dafoam_b200 import cases
from dafoam_b200.pyDASolvers import KSP, Mat, pyDASolvers
from oracle.pyoracle import Oracle
from tests.common import HOSTSIM, NORM_STATES

def setup_case(kind):
    if kind == "naca":
        mesh, bcs, fpatch, ipatch = cases.naca0012_ogrid(ni=24, nj=12, nk=1), cases.default_bcs_naca(), "wing", "inout"
    else:
        mesh, bcs, fpatch, ipatch = cases.channel(nx=16, ny=10, nz=1), cases.default_bcs_channel(), "walls", "inlet"
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": [fpatch], "directionMode": "fixedDirection",
                 "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
    inp = {"patchV": {"type": "patchVelocity", "patches": [ipatch], "flowAxis": "x", "normalAxis": "y", "components": ["solver"]}}
    return mesh, bcs, fn, inp, fpatch, ipatch


def total_derivative(lib_path, kind="channel"):
    mesh, bcs, FN, INPUT, fpatch, ipatch_name = setup_case(kind)
    d = tempfile.mkdtemp(prefix="dab_tot_")
    cases.write_case(d, mesh, bcs)
    opts = dict(normalizeStates=NORM_STATES, function=FN, inputInfo=INPUT,
                adjEqnOption=dict(gmresRelTol=1e-12, gmresMaxIters=800, gmresRestart=800))
    sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=lib_path)
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES)
    n, nC = orc.ndof, mesh.n_cells
    y = np.zeros(nC)
    sol.getOFField("yWall", "scalar", y)
    if kind == "naca":
        W = cases.boundary_layer_state(mesh, y)
    else:
        W = np.zeros(n)
        sol.getOFFields(W)
        W *= 1.0 + 0.02 * np.random.default_rng(3).uniform(-1, 1, n)
    x = np.array([10.0, 3.0])  # |U|, angle of attack [deg]
    a = np.deg2rad(x[1])
    ref = np.array([x[0] * np.cos(a), x[0] * np.sin(a), 0.0])
    # ---- engine: the reference's call order (set input, dFdW, dRdWTPC, KSP solve, dRdx^T psi)
    sol.updateOFFields(W)
    assert sol.getInputSize("patchV", "patchVelocity") == 2
    dFdx = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x, "CD", "function", np.array([1.0]), dFdx)
    dFdW = np.zeros(n)
    sol.calcJacTVecProduct("states", "stateVar", W, "CD", "function", np.array([1.0]), dFdW)
    pc, ksp = Mat(), KSP()
    sol.calcdRdWT(1, pc)
    sol.createMLRKSPMatrixFree(pc, ksp)
    psi = np.zeros(n)
    assert sol.solveLinearEqn(ksp, dFdW, psi) == 0
    dRdxTpsi = np.zeros(2)
    sol.calcJacTVecProduct("patchV", "patchVelocity", x, "R", "residual", psi, dRdxTpsi)
    total = dFdx - dRdxTpsi
    # ---- oracle: tape transposes + dense solve
    ipatch = [p["name"] for p in mesh.patches].index(ipatch_name)
    orc.set_bc_value("U", ipatch, ref)
    orc.record(W)
    A = np.zeros((n, n))
    e = np.zeros(n)
    for i in range(n):
        e[:] = 0.0
        e[i] = 1.0
        A[:, i] = orc.jtvec(e)
    b = orc.dforce_dw(W, [p["name"] for p in mesh.patches].index(fpatch), [1.0, 0.0, 0.0], 1.0)
    psi_o = np.linalg.solve(A, b)
    rb = orc.jtvec_bcU(W, psi_o, ipatch)
    tot_o = -np.array([rb[0] * np.cos(a) + rb[1] * np.sin(a),
                       (-rb[0] * x[0] * np.sin(a) + rb[1] * x[0] * np.cos(a)) * np.pi / 180.0])
    return total, tot_o, psi, psi_o, np.linalg.cond(A)


def check(lib_path):
    # convergent channel (the reference's derivative-test geometry): the tolerance of BASELINE.json's north_star
    total, tot_o, psi, psi_o, cond = total_derivative(lib_path, "channel")
    assert np.linalg.norm(psi - psi_o) <= 1e-6 * np.linalg.norm(psi_o)
    assert np.all(np.abs(tot_o) > 0)
    assert np.all(np.abs(total - tot_o) <= 1e-6 * np.abs(tot_o)), (total, tot_o, cond)
    # NACA0012 O-grid: cell volumes span 8 orders of magnitude and cond(A) ~ 1e10, so a 1e-16 difference in the
    # operator (FMA contraction on the GPU) moves psi by ~cond*eps; the functional is checked to 2e-5
    total, tot_o, psi, psi_o, cond = total_derivative(lib_path, "naca")
    assert np.all(np.abs(total - tot_o) <= 2e-5 * np.abs(tot_o)), (total, tot_o, cond)


def test_total_derivative_host_build():
    check(HOSTSIM)


@pytest.mark.gpu
def test_total_derivative_cuda():
    check(None)


def check_patch_var(lib_path):
    """DAInputPatchVar (scalar and vector boundary values): [dR/dx]^T psi of the engine (central differences on the device
    kernels) against the oracle's tape, incompressible and compressible."""
    from tests.common import setup
    from tests.test_compressible import CONFIGS, setup_comp
    inp = {"nutilda_in": {"type": "patchVar", "varName": "nuTilda", "varType": "scalar", "patches": ["inout"]},
           "U_in": {"type": "patchVar", "varName": "U", "varType": "vector", "patches": ["inout"]},
           "p_out": {"type": "patchVar", "varName": "p", "varType": "scalar", "patches": ["inout"]}}
    for comp in (False, True):
        if comp:
            mesh, orc, sol, W = setup_comp(CONFIGS[0], lib_path)
            sol.updateDAOption(dict(normalizeStates=dict(U=50.0, p=101325.0, T=300.0, nuTilda=1e-3, phi=1.0), inputInfo=inp))
            vals = {"nutilda_in": [6e-5], "U_in": [49.0, 3.0, 0.0], "p_out": [101000.0]}
        else:
            mesh, bcs, orc, sol, W, _ = setup("naca", True, lib_path=lib_path, extra_options=dict(inputInfo=inp))
            vals = {"nutilda_in": [6e-5], "U_in": [9.5, 0.7, 0.0], "p_out": [0.3]}
        sol.updateOFFields(W)
        ip = [p["name"] for p in mesh.patches].index("inout")
        psi = np.random.default_rng(5).uniform(-1, 1, orc.ndof)
        for name, field in (("nutilda_in", "nuTilda"), ("U_in", "U"), ("p_out", "p")):
            x = np.array(vals[name])
            assert sol.getInputSize(name, "patchVar") == len(x)
            prod = np.zeros(len(x))
            sol.calcJacTVecProduct(name, "patchVar", x, "R", "residual", psi, prod)
            orc.set_bc_value(field, ip, x)
            ref = orc.jtvec_bc(W, psi, field, ip)[:len(x)]
            assert np.allclose(prod, ref, rtol=1e-3, atol=1e-6 * np.abs(ref).max()), (comp, name, prod, ref)


def test_patch_var_products_host_build():
    check_patch_var(HOSTSIM)


@pytest.mark.gpu
def test_patch_var_products_cuda():
    check_patch_var(None)
