"""Host readers: ASCII and binary polyMesh give the same engine; dictionaries are parsed like OpenFOAM."""
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
