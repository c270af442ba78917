"""Independent cross-checks of the CPU oracle against scipy / numpy on random inputs ("secondary
cross-check, independent of our restatement").  These do not pin the reference's summation order — they catch a
restatement that computes the wrong THING: (##), the triangular sweeps, the three linSolve0 methods, Arnoldi."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _rand_dd(rng, n, k):
    """diagonally dominant, non-symmetric, k off-diagonals per row"""
    i = np.repeat(np.arange(n), k)
    j = rng.integers(0, n, size=n * k)
    v = rng.uniform(-1, 1, size=n * k)
    A = sp.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A = A.tolil()
    A.setdiag(0.0)
    A = A.tocsr()
    A.eliminate_zeros()
    d = np.asarray(abs(A).sum(axis=1)).ravel() + 1.0
    return (A + sp.diags(d)).tocsr()


def _to_oracle(o, A):
    A = A.tocoo()
    return o.SpMatrix.fromCOO(A.shape, A.row.astype(np.int64), A.col.astype(np.int64), A.data.astype(np.float64))


@pytest.mark.parametrize("seed", range(4))
def test_matmat_vs_scipy(ora, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(40, 30, density=0.2, random_state=seed, format="csr")
    B = sp.random(30, 25, density=0.3, random_state=seed + 100, format="csr")
    C = _to_oracle(ora, A).matMat(_to_oracle(ora, B)).toDense()
    np.testing.assert_allclose(C, (A @ B).toarray(), rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("seed", range(4))
def test_triangular_sweeps_vs_scipy(ora, seed):
    rng = np.random.default_rng(10 + seed)
    n = 60
    A = _rand_dd(rng, n, 5)
    b = rng.uniform(-1, 1, n)
    Ao, bo = _to_oracle(ora, A), ora.SpVector.mkSpVR(n, b)
    w = ora.triLowerSolve(Ao, bo).toDenseListSV()              # reads the lower triangle of a general matrix
    np.testing.assert_allclose(w, spla.spsolve_triangular(sp.tril(A).tocsr(), b, lower=True), rtol=1e-12, atol=1e-14)
    x = ora.triUpperSolve(Ao, bo).toDenseListSV()
    np.testing.assert_allclose(x, spla.spsolve_triangular(sp.triu(A).tocsr(), b, lower=False), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("method", ["BICGSTAB_", "CGS_", "CGNE_"])
def test_linsolve0_vs_dense_solve(ora, method):
    rng = np.random.default_rng(77)
    n = 50
    A = _rand_dd(rng, n, 4)
    xt = rng.uniform(-1, 1, n)
    b = A @ xt
    x, its, res = ora.linSolve0(getattr(ora, method), _to_oracle(ora, A), ora.SpVector.mkSpVR(n, b), ora.SpVector.mkSpVR(n, [0.1] * n), info=True)
    # linSolve0 stops at max(1e-6, 1e-4 * ||r0||) on the TRUE residual (Sparse.hs:1034-1041)
    r0 = np.linalg.norm(b - A @ np.full(n, 0.1))
    assert its < 200 and np.linalg.norm(A @ x.toDenseListSV() - b) <= max(1e-6, 1e-4 * r0) * (1 + 1e-12)
    np.testing.assert_allclose(x.toDenseListSV(), np.linalg.solve(A.toarray(), b), atol=1e-3)


def test_arnoldi_relation_vs_numpy(ora):
    rng = np.random.default_rng(5)
    n, kn = 40, 8
    A = _rand_dd(rng, n, 6)
    b = rng.uniform(-1, 1, n)
    Q, H = ora.arnoldi(_to_oracle(ora, A), ora.SpVector.mkSpVR(n, b), kn)
    Q, H = np.asarray(Q), np.asarray(H)
    assert Q.shape == (n, kn + 1) and H.shape == (kn + 1, kn)
    np.testing.assert_allclose(A @ Q[:, :kn], Q @ H, atol=1e-12)                   # A Q_k = Q_{k+1} H
    np.testing.assert_allclose(Q[:, 0], b / np.linalg.norm(b), atol=1e-15)
    assert np.abs(np.tril(H, -2)).max() == 0.0                                     # upper Hessenberg
