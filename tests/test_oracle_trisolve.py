"""Pins the oracle's preconditioner / triangular-solve restatement against the reference's
own specs (test/LibSpec.hs:203-213, fixtures :1409-1434, checks :436-459) and against values derived by hand from
Sparse.hs:670-811.  CPU only."""
import numpy as np
import pytest

import fixtures as F


@pytest.fixture(scope="module")
def o(ora):
    return ora


# ---- LibSpec.hs:206-213 "triLowerSolve / triUpperSolve (2 x 2 dense, 3 x 3 sparse)": nearZero (norm2 (m #> xhat ^-^ rhs))
@pytest.mark.parametrize("which,mat,rhs", F.TRI_SPECS)
def test_triangular_specs(o, which, mat, rhs):
    m = o.SpMatrix.fromListSM(*mat)
    b = o.SpVector.fromListDenseSV(len(rhs), rhs)
    xhat = o.triLowerSolve(m, b) if which == "lower" else o.triUpperSolve(m, b)
    assert o.nearZero((m.matVec(xhat) - b).norm2())


def test_triangular_hand_derived(o):
    # forward: w0 = 4/2 ; w1 = (10 - 1*2)/4 ; w2 = (17 - 2*2)/3        Sparse.hs:757-766
    w = o.triLowerSolve(o.SpMatrix.fromListSM(*F.LTRI1), o.SpVector.fromListDenseSV(3, F.B_LTRI1)).toDenseListSV()
    assert w.tolist() == [2.0, 2.0, (17.0 - 4.0) / 3.0]
    # backward: x2 = 9/3 ; x1 = (14 - 2*3)/4 ; x0 = (9 - (1*2 + 1*3))/2    Sparse.hs:791-804
    x = o.triUpperSolve(o.SpMatrix.fromListSM(*F.UTRI1), o.SpVector.fromListDenseSV(3, F.B_UTRI1)).toDenseListSV()
    assert x.tolist() == [2.0, 2.0, 3.0]
    # only the named triangle and the diagonal are read: a full matrix gives the same answers
    full = o.SpMatrix.fromListSM((3, 3), F.LTRI1[1] + [(0, 1, 9.0), (0, 2, -4.0), (1, 2, 5.0)])
    assert o.triLowerSolve(full, o.SpVector.fromListDenseSV(3, F.B_LTRI1)).toDenseListSV().tolist() == w.tolist()


def test_triangular_sparsify_and_errors(o):
    # sparsifySV drops |x| <= 1e-12 from the result (Sparse.hs:777) but the sweep itself uses the unsparsified values
    ll = o.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 0, 1e12), (1, 1, 1.0), (2, 2, 1.0)])
    w = o.triLowerSolve(ll, o.SpVector.fromListDenseSV(3, [1e-13, 1.0, 5e-13]))
    assert w.toDenseListSV().tolist() == [0.0, 1.0 - 1e12 * 1e-13, 0.0] and w.nnz == 1
    # NeedsPivoting: a nearZero or missing diagonal, reported for the row the sweep meets first
    with pytest.raises(o.NeedsPivoting) as e:
        o.triLowerSolve(o.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 0, 2.0), (2, 2, 1e-13)]), o.SpVector.fromListDenseSV(3, [1, 1, 1]))
    assert e.value.row == 1
    with pytest.raises(o.NeedsPivoting) as e:
        o.triUpperSolve(o.SpMatrix.fromListSM((3, 3), [(0, 0, 1e-13), (1, 1, 1.0), (2, 2, 1.0)]), o.SpVector.fromListDenseSV(3, [1, 1, 1]))
    assert e.value.row == 0
    # dimension 1: the loop steps before it tests and looks up (1,1) / (-1,-1) with the checked (@@)
    one = o.SpMatrix.fromListSM((1, 1), [(0, 0, 2.0)])
    for f in (o.triLowerSolve, o.triUpperSolve):
        with pytest.raises(o.OracleError) as e:
            f(one, o.SpVector.fromListDenseSV(1, [4.0]))
        assert e.value.code == o.ORA_ERR_OOB_INDEX


@pytest.mark.parametrize("seed", range(4))
def test_prop_triangular_roundtrip(o, seed):
    """L #> x solved for x again (the residual form of LibSpec.hs:436-459 on random triangles)."""
    rng = np.random.default_rng(700 + seed)
    n = int(rng.integers(2, 60))
    d = np.tril(rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.3))
    d[np.arange(n), np.arange(n)] = 2.0 + rng.random(n)
    for tri, solve in ((d, o.triLowerSolve), (d.T.copy(), o.triUpperSolve)):
        i, j = np.nonzero(tri)
        m = o.SpMatrix.fromCOO((n, n), i, j, tri[i, j])
        x = rng.standard_normal(n)
        b = m.matVec(o.SpVector.mkSpVR(n, x))
        xhat = solve(m, b).toDenseListSV()
        np.testing.assert_allclose(xhat, x, rtol=1e-9, atol=1e-11)


def test_diag_partitions_and_preconditioners(o):
    aa = o.SpMatrix.fromListSM((3, 3), [(0, 0, 2), (1, 0, 4), (1, 1, 3), (1, 2, 2), (2, 2, 5), (0, 2, 7)])
    e, d, f = o.diagPartitions(aa)
    assert e.toDense().tolist() == [[0, 0, 0], [4, 0, 0], [0, 0, 0]]
    assert d.toDense().tolist() == [[2, 0, 0], [0, 3, 0], [0, 0, 5]]
    assert f.toDense().tolist() == [[0, 0, 7], [0, 0, 2], [0, 0, 0]]
    assert (e + d + f) == aa                                   # the three parts partition the entries
    # jacobiPre = recip <$> extractDiag                                                  Sparse.hs:686-687
    assert o.jacobiPre(aa).toDense().tolist() == [[0.5, 0, 0], [0, 1 / 3, 0], [0, 0, 0.2]]
    # mSsorPre aa w: l = (I - w e) ## recip d ; r = d - w f                               Sparse.hs:713-721
    w = 1.5
    l, r = o.mSsorPre(aa, w)
    assert l.toDense().tolist() == [[0.5, 0, 0], [0.0 + 0.5 * -(4 * w), 1 / 3, 0], [0, 0, 0.2]]
    assert r.toDense().tolist() == [[2, 0, -(7 * w)], [0, 3, -(2 * w)], [0, 0, 5]]
    assert l.nnz == 9 and r.nnz == 5                           # the (##) stores every (row, column) pair, zeros explicit
    # omega = 1 (symmetric Gauss-Seidel): l = (I - e) d^-1, r = d - f as dense algebra
    l1, r1 = o.mSsorPre(aa, 1.0)
    np.testing.assert_allclose(l1.toDense(), (np.eye(3) - e.toDense()) @ np.diag(1 / np.diag(d.toDense())), atol=1e-15)
    np.testing.assert_allclose(r1.toDense(), d.toDense() - f.toDense(), atol=0)
    # a matrix with a missing diagonal entry: that column disappears from l, jacobiPre has no entry there
    bb = o.SpMatrix.fromListSM((3, 3), [(0, 0, 2), (1, 0, 4), (2, 1, 3), (2, 2, 5)])
    l2, r2 = o.mSsorPre(bb, 2.0)
    assert l2.toDense().tolist() == [[0.5, 0, 0], [0.0 + 0.5 * -(4 * 2.0), 0, 0], [0, 0, 0.2]]
    assert r2.toDense().tolist() == [[2, 0, 0], [0, 0, 0], [0, 0, 5]]
    assert o.jacobiPre(bb).nnz == 2


# ---- lu / ilu0Pre (Sparse.hs:489-538, 696-706): the reference's own specs (LibSpec.hs:184-194, checkLu :424-434) ------------

def _check_lu(o, aa):
    """checkLu: nearZero (normFrobenius (sparsifySM ((l ## u) ^-^ a))) && isUpperTriSM u && isLowerTriSM l"""
    l, u = o.lu(aa)
    L, U, A = l.toDense(), u.toDense(), aa.toDense()
    d = L @ U - A
    d[np.abs(d) <= 1e-12] = 0.0
    return np.sqrt((d * d).sum()) <= 1e-12 and np.allclose(U, np.triu(U)) and np.allclose(L, np.tril(L)), L, U


def test_ref_lu_specs(ora):
    o = ora
    import math

    aa0 = o.SpMatrix.fromListDenseSM(*F.AA0)                                             # "lu (2 x 2 dense)"
    tm0 = o.SpMatrix.fromListSM((2, 2), [(0, 0, math.pi), (1, 0, math.sqrt(2)), (0, 1, math.e), (1, 1, math.sqrt(5))])   # "lu (2 x 2 sparse)"
    tm7 = o.SpMatrix.fromListSM(*F.tm7_triples())                                        # "lu (5 x 5 sparse)"
    for aa in (aa0, tm0, tm7):
        ok, L, U = _check_lu(o, aa)
        assert ok
        assert np.array_equal(np.diag(L), np.ones(L.shape[0]))                           # Doolittle: unit diagonal of L
    # hand-checked: aa0 = [[1,2],[3,4]] -> L = [[1,0],[3,1]], U = [[1,2],[0,-2]]
    _, L, U = _check_lu(o, aa0)
    assert L.tolist() == [[1, 0], [3, 1]] and U.tolist() == [[1, 2], [0, -2]]


def test_lu_against_dense_doolittle_and_pivot_error(ora):
    o = ora
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 30):
        A = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.4) + np.diag(rng.uniform(3, 5, n) * n ** 0.5)
        aa = o.SpMatrix.fromListSM((n, n), [(i, j, A[i, j]) for i in range(n) for j in range(n) if A[i, j] != 0.0])
        l, u = o.lu(aa)
        # independent dense Doolittle (no pivoting), same recurrences in numpy
        L, U = np.eye(n), np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                U[i, j] = A[i, j] - L[i, :i] @ U[:i, j]
            for k in range(i + 1, n):
                L[k, i] = (A[k, i] - L[k, :i] @ U[:i, i]) / U[i, i]
        np.testing.assert_allclose(l.toDense(), L, atol=1e-11)
        np.testing.assert_allclose(u.toDense(), U, atol=1e-11)
        # ilu0Pre: the same numbers where aa stores something, nothing elsewhere (sparsifyLU)
        lh, uh = o.ilu0Pre(aa)
        mask = A != 0.0
        np.testing.assert_array_equal(lh.toDense(), np.where(mask, l.toDense(), 0.0))
        np.testing.assert_array_equal(uh.toDense(), np.where(mask, u.toDense(), 0.0))
        pat = {(i, j) for i in range(n) for j in range(n) if mask[i, j]}
        for m in (lh, uh):
            ii, jj, _ = m.toCOO()
            assert set(zip(ii.tolist(), jj.tolist())) <= pat
    # a zero pivot: u00 = 0 -> NeedsPivoting at (0,0); a pivot that cancels at step 1
    with pytest.raises(o.NeedsPivoting) as e:
        o.lu(o.SpMatrix.fromListSM((2, 2), [(0, 1, 1.0), (1, 0, 1.0)]))
    assert e.value.row == 0
    with pytest.raises(o.NeedsPivoting) as e:
        o.lu(o.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (0, 1, 1.0), (1, 0, 1.0), (1, 1, 1.0), (2, 2, 1.0), (2, 1, 1.0)]))
    assert e.value.row == 1
