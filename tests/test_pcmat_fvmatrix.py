"""calcPCMatWithFvMatrix(PCMat, turbOnly=1) (reference DASolver.C:2888-2988, DASpalartAllmaras.C:490-529; pinned upstream by
tests/runUnitTests_DATurbModel.py:46-53 through the Frobenius norm of the matrix): the engine's matrix against the reference's
insertion loop replayed on the oracle's relaxed nuTilda fvMatrix (D, upper, lower)."""
import numpy as np
import pytest

from tests.common import HOSTSIM, setup


def _reference_insertion(mesh, D, upper, lower, V, sNut, listed, nC):
    """The loops of DASolver::calcPCMatWithFvMatrix, turbulence part: MatSetValues(PCMat, 1, &colI, 1, &rowI, val)."""
    base = 4 * nC
    A = np.zeros((5 * nC + mesh.n_faces,) * 2)
    for c in range(nC):
        A[base + c, base + c] = D[c] * sNut / (V[c] if listed else 1.0)
    for f in range(mesh.n_internal_faces):
        o, n = int(mesh.owner[f]), int(mesh.neighbour[f])
        rowI, colI = base + n, base + o  # "set lower/owner"
        A[colI, rowI] = lower[f] * sNut / (V[n] if listed else 1.0)
        rowI, colI = base + o, base + n  # "set upper/neighbour"
        A[colI, rowI] = upper[f] * sNut / (V[o] if listed else 1.0)
    return A


def _check(lib, kind, nres):
    from dafoam_b200.pyDASolvers import Mat
    mesh, bcs, orc, sol, W, _ = setup(kind, True, nk=1, nres=nres, lib_path=lib)
    sol.updateOFFields(W)
    pc = Mat()
    pc.zeroEntries()
    sol.calcPCMatWithFvMatrix(pc, 1)
    nC = mesh.n_cells
    D, up, lo = orc.nut_fvmatrix(W, alpha=0.7)  # cases.write_case: relaxationFactors nuTilda 0.7
    A = _reference_insertion(mesh, D, up, lo, orc.geometry("V"), 1e-3, "nuTildaRes" in nres, nC)
    B = pc.toDense(orc.ndof)
    assert np.linalg.norm(A) > 0
    assert np.linalg.norm(A - B) <= 1e-11 * np.linalg.norm(A)
    assert abs(pc.norm() - np.linalg.norm(A)) <= 1e-11 * np.linalg.norm(A)  # the number runUnitTests_DATurbModel.py pins
    with pytest.raises(Exception):
        sol.calcPCMatWithFvMatrix(Mat(), 0)  # the reference aborts for this solver family


@pytest.mark.parametrize("kind,nres", [("channel", ("URes", "pRes", "nuTildaRes", "phiRes")), ("naca", ("pRes",)), ("prism", ("URes", "pRes", "nuTildaRes", "phiRes"))])
def test_turbulence_block_from_the_fvmatrix_host_build(kind, nres):
    _check(HOSTSIM, kind, nres)


@pytest.mark.gpu
def test_turbulence_block_from_the_fvmatrix_cuda():
    _check(None, "naca", ("URes", "pRes", "nuTildaRes", "phiRes"))
