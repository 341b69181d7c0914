"""PETSc binary files (big-endian) as the reference writes and reads them through DAUtility::writeMatrixBinary /
readMatrixBinary / writeVectorBinary / readVectorBinary (reference src/adjoint/DAUtility/DAUtility.C:282-441): `dRdWTPC.bin`,
`psi_*.bin`, `dFdW_*.bin`.  Plain numpy: petsc4py is not needed on either side.

Vec:  int32 1211214, int32 n, n float64.
Mat (AIJ): int32 1211216, int32 M, int32 N, int32 nz, M int32 row lengths, nz int32 column indices, nz float64 values."""
import numpy as np

VEC_FILE_CLASSID = 1211214
MAT_FILE_CLASSID = 1211216


def write_vec(path, v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    with open(path, "wb") as f:
        np.array([VEC_FILE_CLASSID, v.size], dtype=">i4").tofile(f)
        v.astype(">f8").tofile(f)


def read_vec(path):
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype=">i4", count=2)
        if head[0] != VEC_FILE_CLASSID:
            raise ValueError("%s is not a PETSc binary Vec" % path)
        return np.fromfile(f, dtype=">f8", count=int(head[1])).astype(np.float64)


def write_mat(path, row_ptr, cols, vals, n_cols=None):
    """CSR (row_ptr[n+1], cols[nnz], vals[nnz]) -> PETSc binary AIJ."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = row_ptr.size - 1
    nnz = int(row_ptr[-1])
    if nnz >= 2**31:
        raise ValueError("more than 2^31 nonzeros: the 32-bit PETSc binary header cannot hold it")
    with open(path, "wb") as f:
        np.array([MAT_FILE_CLASSID, n, n if n_cols is None else n_cols, nnz], dtype=">i4").tofile(f)
        np.diff(row_ptr).astype(">i4").tofile(f)
        np.asarray(cols[:nnz]).astype(">i4").tofile(f)
        np.asarray(vals[:nnz], dtype=np.float64).astype(">f8").tofile(f)


def read_mat(path):
    """-> (n_rows, n_cols, row_ptr, cols, vals)"""
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype=">i4", count=4)
        if head[0] != MAT_FILE_CLASSID:
            raise ValueError("%s is not a PETSc binary Mat" % path)
        m, n, nnz = int(head[1]), int(head[2]), int(head[3])
        lens = np.fromfile(f, dtype=">i4", count=m).astype(np.int64)
        cols = np.fromfile(f, dtype=">i4", count=nnz).astype(np.int32)
        vals = np.fromfile(f, dtype=">f8", count=nnz).astype(np.float64)
    row_ptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    return m, n, row_ptr, cols, vals
