"""Host readers: ASCII and binary polyMesh give the same engine; dictionaries are parsed like OpenFOAM."""
import os

import numpy as np

from tests.common import HOSTSIM, setup


def test_ascii_and_binary_polymesh_agree():
    out = []
    for binary in (False, True):
        mesh, bcs, orc, sol, W, _ = setup("naca", True, nk=1, lib_path=HOSTSIM, binary=binary)
        sol.updateOFFields(W)
        R = np.zeros(orc.ndof)
        sol.getResiduals(R)
        out.append(R)
    assert np.array_equal(out[0], out[1])


def test_initial_states_come_from_the_zero_directory():
    mesh, bcs, orc, sol, W, _ = setup("naca", True, nk=1, lib_path=HOSTSIM)
    W0 = np.zeros(orc.ndof)
    sol.getOFFields(W0)
    nC = mesh.n_cells
    assert np.allclose(W0[:3 * nC].reshape(nC, 3), [10.0, 0.0, 0.0])
    assert np.allclose(W0[4 * nC:5 * nC], 4.5e-5)
    # phi = U_f . Sf: a closed surface integral of a uniform field vanishes cell by cell
    V = orc.geometry("V")
    assert sol.getNLocalCells() == nC and V.min() > 0
    pts = np.zeros(3 * mesh.n_points)
    sol.getOFMeshPoints(pts)
    assert np.array_equal(pts.reshape(-1, 3), mesh.points)


def test_field_output_round_trip():
    """writeFields / writeAdjointFields (OpenFOAM ASCII vol and surface fields) and reading the written states back as the 0/
    fields of a case: an exact restart."""
    import re
    import shutil
    import tempfile
    from dafoam_b200 import cases
    from dafoam_b200.pyDASolvers import pyDASolvers
    from tests.common import HOSTSIM
    mesh, bcs = cases.channel(nx=10, ny=6, nz=1), cases.default_bcs_channel()
    d = tempfile.mkdtemp(prefix="dab_out_")
    cases.write_case(d, mesh, bcs)
    sol = pyDASolvers("DASimpleFoam -python", dict(primalMaxIters=25), caseDir=d, _lib_path=HOSTSIM)
    sol.solvePrimal()
    n = sol.getNLocalAdjointStates()
    W = np.zeros(n)
    sol.getOFFields(W)
    sol.writeFields(25)
    psi = np.random.default_rng(2).uniform(-1, 1, n)
    sol.writeAdjointFields("CD", 25, psi)
    for name in ("U", "p", "nuTilda", "phi", "adjoint_CD_U", "adjoint_CD_p", "adjoint_CD_nuTilda", "adjoint_CD_phi"):
        assert os.path.exists(os.path.join(d, "25", name)), name
    txt = open(os.path.join(d, "25", "adjoint_CD_p")).read()
    vals = np.array(re.search(r"internalField nonuniform List<scalar> \d+\s*\(([^)]*)\)", txt).group(1).split(), dtype=float)
    nC = mesh.n_cells
    assert np.array_equal(vals, psi[3 * nC:4 * nC])
    # restart: the case's own boundary conditions with the written internal fields (+ the written flux)
    d2 = tempfile.mkdtemp(prefix="dab_restart_")
    shutil.copytree(d, d2, dirs_exist_ok=True)
    for name in ("U", "p", "nuTilda"):
        new = re.search(r"internalField[^;]*;", open(os.path.join(d, "25", name)).read(), re.S).group(0)
        old = open(os.path.join(d2, "0", name)).read()
        open(os.path.join(d2, "0", name), "w").write(re.sub(r"internalField[^;]*;", lambda m: new, old, count=1, flags=re.S))
    shutil.copy(os.path.join(d, "25", "phi"), os.path.join(d2, "0", "phi"))
    sol2 = pyDASolvers("DASimpleFoam -python", {}, caseDir=d2, _lib_path=HOSTSIM)
    W2 = np.zeros(n)
    sol2.getOFFields(W2)
    assert np.array_equal(W, W2)


def test_pc_matrix_export_petsc_binary():
    """writeJacobians: ["dRdWTPC"] leaves dRdWTPC.bin (PETSc binary AIJ, reference DAUtility::writeMatrixBinary) in the case
    directory: the matrix is the coloured-FD transpose Jacobian of the first-order residual, checked column by column against
    finite differences of getResiduals(isPC=1)."""
    from dafoam_b200 import petsc_io
    from dafoam_b200.pyDASolvers import Mat
    from tests.common import NORM_STATES
    mesh, bcs, orc, sol, W, _ = setup("naca", True, nk=1, lib_path=HOSTSIM, extra_options=dict(writeJacobians=["dRdWTPC"]))
    sol.updateOFFields(W)
    sol.calcdRdWT(1, Mat())
    path = os.path.join(sol._caseDir, "dRdWTPC.bin")
    m, n, rp, cl, vl = petsc_io.read_mat(path)
    assert m == n == orc.ndof and rp[-1] == len(cl) == len(vl)
    assert all(np.all(np.diff(cl[rp[i]:rp[i + 1]]) > 0) for i in range(0, m, 97))  # sorted columns
    nC = mesh.n_cells
    scale = np.concatenate([np.full(3 * nC, NORM_STATES["U"]), np.full(nC, NORM_STATES["p"]), np.full(nC, NORM_STATES["nuTilda"]),
                            NORM_STATES["phi"] * orc.geometry("magSf")])
    R0 = np.zeros(m)
    sol.getResiduals(R0, 1)
    rng = np.random.default_rng(0)
    for i in rng.choice(m, 12, replace=False):
        Wp = W.copy()
        h = 1e-6 * scale[i]
        Wp[i] += h
        sol.updateOFFields(Wp)
        R1 = np.zeros(m)
        sol.getResiduals(R1, 1)
        col = (R1 - R0) / 1e-6  # row i of dRdWT in the scaled-state convention of the reference (perturbation = eps * scale)
        row = np.zeros(m)
        row[cl[rp[i]:rp[i + 1]]] = vl[rp[i]:rp[i + 1]]
        inside = np.zeros(m, dtype=bool)
        inside[cl[rp[i]:rp[i + 1]]] = True
        assert np.abs(row - col)[inside].max() <= 1e-4 * max(np.abs(col).max(), 1e-30), i
    sol.updateOFFields(W)
    # the vector format round-trips
    petsc_io.write_vec(os.path.join(sol._caseDir, "psi.bin"), W)
    assert np.array_equal(petsc_io.read_vec(os.path.join(sol._caseDir, "psi.bin")), W)


def test_read_state_vars_mesh_points_and_sens_maps():
    """readStateVars / readMeshPoints / writeMeshPoints / writeSensMapField / writeSensMapSurface of the reference's pyDASolvers
    (pyDASolvers.pyx:382-395, 421-462): a written time directory is read back exactly; the surface map follows DASolver.C:3840-3960."""
    import re
    import tempfile
    from dafoam_b200 import cases
    from dafoam_b200.pyDASolvers import pyDASolvers
    mesh, bcs = cases.naca0012_ogrid(ni=24, nj=12, nk=1), cases.default_bcs_naca()
    d = tempfile.mkdtemp(prefix="dab_rs_")
    cases.write_case(d, mesh, bcs)
    sol = pyDASolvers("DASimpleFoam -python", dict(primalMaxIters=10, inputInfo={"shape": {"type": "volCoord", "components": ["solver"]}}),
                      caseDir=d, _lib_path=HOSTSIM)
    assert sol.hasVolCoordInput() == 1 and sol.getGlobalXvIndex(3, 2) == 11
    n, nC, nP = sol.getNLocalAdjointStates(), sol.getNLocalCells(), sol.getNLocalPoints()
    assert sol.getNLocalAdjointBoundaryStates() == 5 * (sol.getNLocalFaces() - sol.getNLocalInternalFaces())
    W0 = np.zeros(n)
    sol.getOFFields(W0)
    sol.solvePrimal()
    assert sol.getPrevPrimalSolTime() == 10.0 and sol.getDeltaT() == 1.0 and sol.getDdtSchemeOrder() == 1
    W = np.zeros(n)
    sol.getOFFields(W)
    sol.writeFields(10)
    assert sol.getLatestTime() == 10.0
    # move the mesh, write it under the same time, restore, read both back
    pts0 = np.zeros(3 * nP)
    sol.getOFMeshPoints(pts0)
    pts1 = pts0 + 1e-3 * np.sin(np.arange(3 * nP))
    sol.writeMeshPoints(pts1, 10)
    sol.updateOFFields(W0)
    sol.readStateVars(10, 0)
    W2 = np.zeros(n)
    sol.getOFFields(W2)
    assert np.array_equal(W, W2)
    sol.readMeshPoints(10)
    pts2 = np.zeros(3 * nP)
    sol.getOFMeshPoints(pts2)
    assert np.array_equal(pts1, pts2)
    sol.updateOFMesh(pts0)
    sol.writeCurrentMeshPointsToConstant()
    try:
        sol.readStateVars(11, 0)
        assert False
    except Exception as e:
        assert "does not exist" in str(e)
    # sensitivity maps
    g = np.random.default_rng(5).normal(size=3 * nC)
    sol.writeSensMapField("dFdU", g, "vector", 10)
    txt = open(os.path.join(d, "10", "dFdU")).read()
    body = re.search(r"internalField nonuniform List<vector> \d+\s*\((.*?)\n\);", txt, re.S).group(1)
    assert np.array_equal(np.array(body.replace("(", " ").replace(")", " ").split(), dtype=float), g)
    sol.writeSensMapField("dFdp", g[:nC], "scalar", 10)
    assert "volScalarField" in open(os.path.join(d, "10", "dFdp")).read()
    # surface map: the design surface = the wall points themselves, derivative = a function of the point -> each wall face holds
    # the sum over its 4 points divided by 3, and the closest-distance norm is zero
    wall = [p for p in mesh.patches if p["name"] == "wing"][0]
    faces = [mesh.faces[wall["start"] + i] for i in range(wall["size"])]
    ids = sorted({int(v) for f in faces for v in f})
    Xs = pts0.reshape(-1, 3)[ids].copy()
    dF = np.stack([Xs[:, 0] + 2.0, 3.0 * Xs[:, 1], np.ones(len(ids))], axis=1)
    nrm = sol.writeSensMapSurface("sensCD", dF.ravel().copy(), Xs.ravel().copy(), 3 * len(ids), 10)
    assert nrm == 0.0
    txt = open(os.path.join(d, "10", "sensCD")).read()
    blk = re.search(r"wing\s*\{[^}]*?value nonuniform List<vector> (\d+)\((.*?)\);", txt, re.S)
    assert int(blk.group(1)) == wall["size"]
    got = np.array(blk.group(2).replace("(", " ").replace(")", " ").split(), dtype=float).reshape(-1, 3)
    pos = {v: i for i, v in enumerate(ids)}
    want = np.array([sum(dF[pos[int(v)]] for v in f) / 3.0 for f in faces])
    assert np.allclose(got, want, rtol=1e-14, atol=0)
    assert sol.getElapsedClockTime() > 0 and sol.getElapsedCpuTime() > 0
    g = np.full(sol.getNGlobalCells(), -1.0)
    sol.getOFFieldGlobal("p", "scalar", g)
    assert np.array_equal(g, W[3 * nC:4 * nC])
    sol.setPrimalBoundaryConditions(0)
    iv = sol.getInitStateVals(0)
    assert set(iv) == {"U0", "U1", "U2", "p", "nuTilda"} and abs(iv["U0"] - W[0:3 * nC:3].mean()) < 1e-12 and abs(iv["nuTilda"] - W[4 * nC:5 * nC].mean()) < 1e-15


def test_check_mesh():
    """checkMesh (DACheckMesh): the O-grid passes; a point pushed through its neighbours gives negative pyramids / volumes and fails."""
    import tempfile
    from dafoam_b200 import cases
    from dafoam_b200.pyDASolvers import pyDASolvers
    mesh, bcs = cases.naca0012_ogrid(ni=24, nj=12, nk=1), cases.default_bcs_naca()
    d = tempfile.mkdtemp(prefix="dab_cm_")
    cases.write_case(d, mesh, bcs)
    # this coarse O-grid has two skew trailing-edge faces (4.06 > the default 4): fails by default, passes with a looser threshold
    sol0 = pyDASolvers("DASimpleFoam -python", {}, caseDir=d, _lib_path=HOSTSIM)
    assert sol0.checkMesh() == 0 and sol0.meshQuality["nFailedChecks"] == 1 and 4.0 < sol0.meshQuality["maxSkewness"] < 4.2
    sol = pyDASolvers("DASimpleFoam -python", dict(checkMeshThreshold=dict(maxSkewness=6.0)), caseDir=d, _lib_path=HOSTSIM)
    assert sol.checkMesh() == 1
    q = sol.meshQuality
    assert 0 <= q["maxNonOrth"] < 75 and q["nNegativePyramids"] == 0 and q["minVolume"] > 0 and q["maxOpenness"] < 1e-12
    # the convergent channel passes with the defaults
    dc = tempfile.mkdtemp(prefix="dab_cmc_")
    cases.write_case(dc, cases.channel(nx=10, ny=6, nz=1), cases.default_bcs_channel())
    solc = pyDASolvers("DASimpleFoam -python", {}, caseDir=dc, _lib_path=HOSTSIM)
    assert solc.checkMesh() == 1 and solc.meshQuality["maxNonOrth"] < 70 and solc.meshQuality["nSevereNonOrth"] == 0
    # tighter thresholds than the mesh meets: aspect ratio / skewness become failures
    sol2 = pyDASolvers("DASimpleFoam -python", dict(checkMeshThreshold=dict(maxAspectRatio=1.0 + 1e-9, maxSkewness=2.0)),
                       caseDir=d, _lib_path=HOSTSIM)
    assert sol2.checkMesh() == 0 and sol2.meshQuality["nFailedChecks"] == 2
    # tangle the mesh: push an interior point (both z-copies) 1.3 edge lengths past its closest neighbour
    pts = np.zeros(3 * sol.getNLocalPoints())
    sol.getOFMeshPoints(pts)
    P = pts.reshape(-1, 3)
    r = np.hypot(P[:, 0] - 0.5, P[:, 1])
    i = int(np.argsort(r)[len(r) // 2])
    same = np.where((np.abs(P[:, 0] - P[i, 0]) < 1e-12) & (np.abs(P[:, 1] - P[i, 1]) < 1e-12))[0]
    dd = np.hypot(P[:, 0] - P[i, 0], P[:, 1] - P[i, 1])
    dd[same] = 1e9
    j = int(np.argmin(dd))
    P[same, :2] += 1.3 * (P[j, :2] - P[i, :2])
    sol.updateOFMesh(pts)
    assert sol.checkMesh() == 0 and sol.meshQuality["nNegativePyramids"] > 0 and sol.meshQuality["nErrorNonOrth"] > 0


def test_solve_nonlinear_refuses_a_failed_mesh():
    """PYDAFOAM.solve_nonlinear (reference mphys_dafoam.py:314-368): a mesh that fails checkMesh is written for inspection and the
    primal is not run; a good mesh returns the converged states; prepareCaseOnly only writes the points to constant/."""
    import tempfile
    import pytest
    from dafoam_b200 import cases
    from dafoam_b200.pyDAFoam import PYDAFOAM, AnalysisError
    d = tempfile.mkdtemp(prefix="dab_sn_")
    cases.write_case(d, cases.channel(nx=10, ny=6, nz=1), cases.default_bcs_channel())
    opts = dict(solverName="DASimpleFoam", primalMaxIters=5, primalMinResTol=1e-30, primalMinResTolDiff=1e30)
    DASolver = PYDAFOAM(options=opts, caseDir=d, _lib_path=HOSTSIM)
    W = DASolver.solve_nonlinear()
    assert W.shape == (DASolver.getNLocalAdjointStates(),) and DASolver.nSolvePrimals == 2
    DASolver.setOption("checkMeshThreshold", dict(maxAspectRatio=1.01))
    DASolver.solver._options["checkMeshThreshold"] = dict(maxAspectRatio=1.01)
    with pytest.raises(AnalysisError, match="Mesh quality"):
        DASolver.solve_nonlinear()
    assert DASolver.nSolvePrimals == 2 and os.path.exists(os.path.join(d, "9999", "polyMesh", "points"))
    DASolver.solver._options["checkMeshThreshold"] = {}
    DASolver.setOption("prepareCaseOnly", True)
    assert DASolver.solve_nonlinear() is None and DASolver.nSolvePrimals == 2
