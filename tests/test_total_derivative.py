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

FN = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
             "direction": [1.0, 0.0, 0.0], "scale": 1.0}}
INPUT = {"patchV": {"type": "patchVelocity", "patches": ["inout"], "flowAxis": "x", "normalAxis": "y", "components": ["solver"]}}


def total_derivative(lib_path):
    mesh = cases.naca0012_ogrid(ni=24, nj=12, nk=1)
    bcs = cases.default_bcs_naca()
    d = tempfile.mkdtemp(prefix="dab_tot_")
    cases.write_case(d, mesh, bcs)
    opts = dict(normalizeStates=NORM_STATES, function=FN, inputInfo=INPUT,
                adjEqnOption=dict(gmresRelTol=1e-11, gmresMaxIters=600, gmresRestart=600))
    sol = pyDASolvers("DASimpleFoam -python", opts, caseDir=d, _lib_path=lib_path)
    orc = Oracle(mesh, bcs, normalizeStates=NORM_STATES)
    n, nC = orc.ndof, mesh.n_cells
    y = np.zeros(nC)
    sol.getOFField("yWall", "scalar", y)
    W = cases.boundary_layer_state(mesh, y)
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
    ipatch = [p["name"] for p in mesh.patches].index("inout")
    orc.set_bc_value("U", ipatch, ref)
    orc.record(W)
    A = np.zeros((n, n))
    e = np.zeros(n)
    for i in range(n):
        e[:] = 0.0
        e[i] = 1.0
        A[:, i] = orc.jtvec(e)
    b = orc.dforce_dw(W, 0, [1.0, 0.0, 0.0], 1.0)
    psi_o = np.linalg.solve(A, b)
    rb = orc.jtvec_bcU(W, psi_o, ipatch)
    tot_o = -np.array([rb[0] * np.cos(a) + rb[1] * np.sin(a),
                       (-rb[0] * x[0] * np.sin(a) + rb[1] * x[0] * np.cos(a)) * np.pi / 180.0])
    return total, tot_o, psi, psi_o


def check(lib_path):
    total, tot_o, psi, psi_o = total_derivative(lib_path)
    assert np.linalg.norm(psi - psi_o) <= 1e-6 * np.linalg.norm(psi_o)
    assert np.all(np.abs(tot_o) > 0)
    assert np.all(np.abs(total - tot_o) <= 1e-6 * np.abs(tot_o)), (total, tot_o)


def test_total_derivative_host_build():
    check(HOSTSIM)


@pytest.mark.gpu
def test_total_derivative_cuda():
    check(None)
