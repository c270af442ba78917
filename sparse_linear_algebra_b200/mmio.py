"""Matrix Market ingest (host side): the on-disk format the reference ships its real-matrix fixtures in
(test/data/e05r0000.mtx, read by test/Perf.hs:14-45 through Data.Matrix.MatrixMarket).  Coordinate files are
1-based (Perf.hs:36-37 subtracts 1); `general` and `symmetric` real matrices and dense `array` vectors."""
import numpy as np


def _header(f):
    first = f.readline().strip().split()
    if len(first) < 5 or first[0].lower() != "%%matrixmarket":
        raise ValueError("not a Matrix Market file")
    obj, fmt, field, symm = (t.lower() for t in first[1:5])
    line = f.readline()
    while line.startswith("%") or not line.strip():
        line = f.readline()
    return obj, fmt, field, symm, line.split()


def read_matrix_market(path):
    """Coordinate real matrix -> (m, n, i, j, v) with 0-based int64 indices, in file order."""
    with open(path) as f:
        obj, fmt, field, symm, size = _header(f)
        if obj != "matrix" or fmt != "coordinate" or field not in ("real", "integer", "double"):
            raise ValueError(f"unsupported Matrix Market kind: {obj} {fmt} {field}")
        m, n, nnz = int(size[0]), int(size[1]), int(size[2])
        data = np.loadtxt(f, dtype=np.float64, ndmin=2) if nnz else np.zeros((0, 3))
    if data.shape[0] != nnz:
        raise ValueError(f"expected {nnz} entries, found {data.shape[0]}")
    i = data[:, 0].astype(np.int64) - 1
    j = data[:, 1].astype(np.int64) - 1
    v = data[:, 2].copy()
    if symm == "symmetric":
        off = i != j
        i, j, v = np.concatenate([i, j[off]]), np.concatenate([j, i[off]]), np.concatenate([v, v[off]])
    elif symm != "general":
        raise ValueError(f"unsupported symmetry: {symm}")
    return m, n, i, j, v


def read_array(path):
    """Dense `array` file -> column-major values reshaped to (rows, cols)."""
    with open(path) as f:
        obj, fmt, field, symm, size = _header(f)
        if fmt != "array":
            raise ValueError("not an array file")
        rows, cols = int(size[0]), int(size[1])
        vals = np.loadtxt(f, dtype=np.float64).reshape(-1)
    return vals.reshape(cols, rows).T.copy()
