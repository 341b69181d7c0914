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
