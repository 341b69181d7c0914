"""The `volCoord` input (reference DAInputVolCoord.C through calcJacTVecProduct): [dR/dx_v]^T psi and dF/dx_v of the engine
(coloured central differences over the mesh points on the device) against the oracle's exact tape product through the
templated geometry pipeline, and updateOFMesh."""
import numpy as np
import pytest

from tests.common import HOSTSIM, rel_err, setup


def run_residual_product(lib_path, kind, tol, nk=1):
    mesh, bcs, orc, sol, W, _ = setup(kind, True, "linearUpwind", nk, lib_path=lib_path)
    sol.updateOFFields(W)
    nP3 = 3 * sol.getNLocalPoints()
    assert sol.getInputSize("aero_vol_coords", "volCoord") == nP3
    pts = np.zeros(nP3)
    sol.getOFMeshPoints(pts)
    rng = np.random.default_rng(7)
    psi = rng.uniform(-1, 1, orc.ndof)
    prod = np.zeros(nP3)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "R", "residual", psi, prod)
    ref = orc.jtvec_xv(W, psi)
    # Points on a symmetry plane: the derivative ALONG THE PLANE NORMAL differentiates |n_k| of the symmetry transform
    # coefficients at n_k = 0, a kink -- the tape takes the one-sided CoDiPack convention (fabs'(0) = +1), central
    # differences its symmetric value.  Those components are constrained in practice (symmetry-plane points move in the
    # plane); every other component must agree.
    mask = np.ones((len(pts) // 3, 3), dtype=bool)
    for pch in mesh.patches:
        if pch["type"] == "symmetry":
            fp = mesh.faces[pch["start"]:pch["start"] + pch["size"]]
            mask[np.unique(fp[fp >= 0]), 2] = False  # the symmetry planes of the synthetic cases are z = const
    mask = mask.ravel()
    err = rel_err(prod[mask], ref[mask])
    assert err < tol, err
    assert rel_err(prod[~mask], ref[~mask]) < 0.05  # same magnitude, convention-dependent
    # the unperturbed geometry is restored: the residual is bitwise what it was
    R0, R1 = np.zeros(orc.ndof), np.zeros(orc.ndof)
    sol.getResiduals(R1)
    sol.updateOFFields(W)
    sol.getResiduals(R0)
    assert np.array_equal(R0, R1)
    return err


def run_function_and_mesh_update(lib_path):
    dirv = [float(np.cos(0.05)), float(np.sin(0.05)), 0.0]
    fn = {"CD": {"type": "force", "source": "patchToFace", "patches": ["wing"], "directionMode": "fixedDirection",
                 "direction": dirv, "scale": 0.02}}
    mesh, bcs, orc, sol, W, _ = setup("naca", True, "linearUpwind", 1, lib_path=lib_path, extra_options=dict(function=fn))
    sol.updateOFFields(W)
    nP3 = 3 * sol.getNLocalPoints()
    pts = np.zeros(nP3)
    sol.getOFMeshPoints(pts)
    dFdx = np.zeros(nP3)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "CD", "function", np.array([1.0]), dFdx)
    assert np.linalg.norm(dFdx) > 0
    # directional check through updateOFMesh: F(x + h v) - F(x - h v) = 2 h dFdx . v for a smooth displacement field v
    X = pts.reshape(-1, 3)
    v = np.stack([np.sin(3.0 * X[:, 1]) * 1e-3, np.cos(2.0 * X[:, 0]) * 1e-3, np.zeros(len(X))], axis=1).ravel()
    h = 1e-3
    sol.updateOFMesh(pts + h * v)
    Fp = sol.calcFunction("CD")
    sol.updateOFMesh(pts - h * v)
    Fm = sol.calcFunction("CD")
    sol.updateOFMesh(pts)
    fd = (Fp - Fm) / (2 * h)
    assert abs(fd - dFdx @ v) <= 1e-5 * abs(fd), (fd, dFdx @ v)
    # ... and the same for psi . R
    psi = np.random.default_rng(3).uniform(-1, 1, orc.ndof)
    prod = np.zeros(nP3)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "R", "residual", psi, prod)
    Rp, Rm = np.zeros(orc.ndof), np.zeros(orc.ndof)
    sol.setSolverInput("aero_vol_coords", "volCoord", nP3, pts + h * v)
    sol.getResiduals(Rp)
    sol.setSolverInput("aero_vol_coords", "volCoord", nP3, pts - h * v)
    sol.getResiduals(Rm)
    sol.updateOFMesh(pts)
    fd = psi @ (Rp - Rm) / (2 * h)
    assert abs(fd - prod @ v) <= 1e-5 * abs(fd), (fd, prod @ v)


def test_volcoord_residual_product_matches_oracle_tape_host_build():
    for kind, nk in (("naca", 1), ("channel", 2), ("prism", 1)):
        err = run_residual_product(HOSTSIM, kind, 1e-7, nk)
        print(kind, "volCoord product vs oracle tape: %.2e" % err)


def run_compressible(lib_path):
    """DARhoSimpleFoam: the same coloured-FD product on the 6-state layout vs the oracle's tape through the geometry."""
    from tests.test_compressible import CONFIGS, setup_comp
    cfg = ("naca", "sensibleInternalEnergy", "const", "SpalartAllmaras", "linearUpwind", "upwind", False, CONFIGS[0][7])
    mesh, orc, sol, W = setup_comp(cfg, lib_path)
    sol.updateOFFields(W)
    nP3 = 3 * sol.getNLocalPoints()
    pts = np.zeros(nP3)
    sol.getOFMeshPoints(pts)
    psi = np.random.default_rng(7).uniform(-1, 1, orc.ndof)
    prod = np.zeros(nP3)
    sol.calcJacTVecProduct("aero_vol_coords", "volCoord", pts, "R", "residual", psi, prod)
    ref = orc.jtvec_xv(W, psi)
    mask = np.ones((nP3 // 3, 3), dtype=bool)
    for pch in mesh.patches:
        if pch["type"] == "symmetry":
            fp = mesh.faces[pch["start"]:pch["start"] + pch["size"]]
            mask[np.unique(fp[fp >= 0]), 2] = False
    mask = mask.ravel()
    assert rel_err(prod[mask], ref[mask]) < 1e-7


def test_volcoord_compressible_host_build():
    run_compressible(HOSTSIM)


@pytest.mark.gpu
def test_volcoord_compressible_cuda():
    run_compressible(None)


def test_volcoord_function_and_mesh_update_host_build():
    run_function_and_mesh_update(HOSTSIM)


@pytest.mark.gpu
def test_volcoord_residual_product_matches_oracle_tape_cuda():
    assert run_residual_product(None, "naca", 1e-6) < 1e-6


@pytest.mark.gpu
def test_volcoord_function_and_mesh_update_cuda():
    run_function_and_mesh_update(None)
