"""GPU parity tests for the preconditioners and sparse triangular solves, through the C ABI,
against the CPU oracle and the reference's own fixtures (test/LibSpec.hs:203-213, :1409-1434).

Tolerances: everything here is BIT-EXACT — the partitions and preconditioners are copies / single roundings, the
triangular sweeps evaluate r = sum l_ij w_j in ascending j with one rounding per product and per addition, exactly
the reference's left fold (Sparse.hs:762, 795), followed by one subtraction and one division.
"""
import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sla():
    import sparse_linear_algebra_b200 as s

    return s


@pytest.fixture(scope="module")
def o(ora):
    return ora


def both(sla, o, n, i, j, v):
    return sla.SpMatrix.fromCOO((n, n), i, j, v), o.SpMatrix.fromCOO((n, n), i, j, v)


def same_bits(a, b):
    return np.asarray(a, dtype=np.float64).tobytes() == np.asarray(b, dtype=np.float64).tobytes()


def check_solves(sla, o, A, Ao, b, which=("lower", "upper")):
    n = len(b)
    bd, bo = sla.SpVector.mkSpVR(n, b), o.SpVector.mkSpVR(n, b)
    for w in which:
        got = (sla.triLowerSolve if w == "lower" else sla.triUpperSolve)(A, bd).toDenseListSV()
        want = (o.triLowerSolve if w == "lower" else o.triUpperSolve)(Ao, bo).toDenseListSV()
        assert same_bits(got, want), f"{w}: max |d| = {np.nanmax(np.abs(got - want))}"


# =============================================================== the reference's own specs, on the device

@pytest.mark.parametrize("which,mat,rhs", F.TRI_SPECS)
def test_ref_triangular_specs(sla, which, mat, rhs):    # LibSpec.hs:206-213, checks :436-459
    m = sla.SpMatrix.fromListSM(*mat)
    b = sla.SpVector.fromListDenseSV(len(rhs), rhs)
    xhat = sla.triLowerSolve(m, b) if which == "lower" else sla.triUpperSolve(m, b)
    assert abs(((m @ xhat) - b).norm2()) <= 1e-12
    assert xhat.toDenseListSV().tolist()[:2] == [2.0, 2.0 if len(rhs) == 3 else 3.0]


# =============================================================== bit-exact against the oracle

@pytest.mark.parametrize("kind,n,k,band", [("uniform", 3000, 8, 0), ("uniform", 20000, 16, 0), ("banded", 6000, 12, 40),
                                           ("laplace", 64 * 64, 5, 64), ("laplace", 37 * 37, 5, 37)])
def test_sweeps_on_general_matrices(sla, o, kind, n, k, band):
    """A general (diagonally dominant) matrix: the forward sweep reads its lower triangle, the backward sweep its
    upper triangle (Gauss-Seidel sweeps).  Level counts: the 5-point stencil has 2 * side - 1 wavefronts."""
    gk = {"uniform": (sla.GEN_UNIFORM, o.GEN_UNIFORM), "banded": (sla.GEN_BANDED, o.GEN_BANDED), "laplace": (sla.GEN_LAPLACE2D, o.GEN_LAPLACE2D)}[kind]
    seed = 0x5EED0010
    A = sla.SpMatrix.generate(gk[0], n, k, seed, band)
    Ao = o.SpMatrix.synth(gk[1], n, k, seed, band)
    b = o.SpVector.synth(seed + 1, n).toDenseListSV()
    check_solves(sla, o, A, Ao, b)
    if kind == "laplace":
        assert A.triAnalysis(False)[0] == 2 * band - 1 and A.triAnalysis(True)[0] == 2 * band - 1
        assert A.triAnalysis(False)[1] == n + 2 * band * (band - 1)      # diagonal + west + north neighbours
    # the schedule is cached: a second solve with another right-hand side
    check_solves(sla, o, A, Ao, b[::-1].copy())


def test_chain_and_dense_triangles(sla, o):
    """Worst and best cases of the schedule: a bidiagonal chain (n levels, every row waits for its neighbour, so lanes
    of one warp depend on each other), a diagonal (1 level), a dense triangle (long rows, n levels)."""
    rng = np.random.default_rng(71)
    n = 3000
    i = np.concatenate([np.arange(n), np.arange(1, n)])
    j = np.concatenate([np.arange(n), np.arange(0, n - 1)])
    v = np.concatenate([1.5 + rng.random(n), 0.5 * rng.standard_normal(n - 1)])
    A, Ao = both(sla, o, n, i, j, v)
    b = rng.standard_normal(n)
    check_solves(sla, o, A, Ao, b, ("lower",))
    assert A.triAnalysis(False) == (n, 2 * n - 1)
    assert A.triAnalysis(True) == (1, n)
    At, Aot = both(sla, o, n, j, i, v)                                     # the same chain, upper bidiagonal
    check_solves(sla, o, At, Aot, b, ("upper",))
    assert At.triAnalysis(True)[0] == n
    m = 300
    d = np.tril(rng.standard_normal((m, m)))
    d[np.arange(m), np.arange(m)] = 20.0 + rng.random(m)
    ii, jj = np.nonzero(d)
    L, Lo = both(sla, o, m, ii, jj, d[ii, jj])
    check_solves(sla, o, L, Lo, rng.standard_normal(m), ("lower",))
    U, Uo = both(sla, o, m, jj, ii, d[ii, jj])
    check_solves(sla, o, U, Uo, rng.standard_normal(m), ("upper",))
    assert L.triAnalysis(False)[0] == m and U.triAnalysis(True)[0] == m


def test_random_triangles_with_ragged_rows(sla, o):
    """Random sparsity incl. rows that only hold their diagonal, and values spanning many magnitudes."""
    rng = np.random.default_rng(72)
    n = 5000
    dens = rng.random(n) * 0.004
    rows, cols, vals = [np.arange(n)], [np.arange(n)], [np.where(rng.random(n) < 0.5, 1.0, -1.0) * (1.0 + rng.random(n))]
    for r in range(1, n, 1):
        cnt = rng.binomial(r, dens[r]) if r % 7 else 0
        if cnt:
            c = rng.choice(r, size=cnt, replace=False)
            rows.append(np.full(cnt, r)); cols.append(c); vals.append(rng.standard_normal(cnt) * 10.0 ** rng.integers(-3, 2, cnt))
    i, j, v = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    b = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
    A, Ao = both(sla, o, n, i, j, v)
    check_solves(sla, o, A, Ao, b, ("lower",))
    At, Aot = both(sla, o, n, j, i, v)
    check_solves(sla, o, At, Aot, b, ("upper",))


def test_sparsify_and_special_values(sla, o):
    """sparsifySV on the result only (Sparse.hs:777, 811); infinities and NaNs travel like in the oracle."""
    i, j, v = [0, 1, 1, 2, 3, 3], [0, 0, 1, 2, 0, 3], [1.0, 1e12, 1.0, 1.0, 1.0, 1.0]
    A, Ao = both(sla, o, 4, i, j, v)
    b = np.array([1e-13, 1.0, 5e-13, 0.0])
    got = sla.triLowerSolve(A, sla.SpVector.mkSpVR(4, b)).toDenseListSV()
    want = o.triLowerSolve(Ao, o.SpVector.mkSpVR(4, b)).toDenseListSV()
    assert same_bits(got, want) and got[0] == 0.0 and got[2] == 0.0 and got[1] == 1.0 - 1e12 * 1e-13 and got[3] == 0.0
    b2 = np.array([np.inf, 1.0, np.nan, 2.0])
    got = sla.triLowerSolve(A, sla.SpVector.mkSpVR(4, b2)).toDenseListSV()
    want = o.triLowerSolve(Ao, o.SpVector.mkSpVR(4, b2)).toDenseListSV()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_needs_pivoting_and_dimension_errors(sla, o):
    one = sla.SpVector.mkSpVR(3, [1.0, 1.0, 1.0])
    with pytest.raises(sla.NeedsPivoting) as e:       # missing diagonal in row 1, nearZero in row 2: the forward sweep meets 1 first
        sla.triLowerSolve(sla.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 0, 2.0), (2, 2, 1e-13)]), one)
    assert "L (1,1)" in e.value.message
    with pytest.raises(sla.NeedsPivoting) as e:       # the backward sweep meets row 2 first and reports it as (0,0), Sparse.hs:802
        sla.triUpperSolve(sla.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 0, 2.0), (2, 2, 1e-13)]), one)
    assert "U (0,0)" in e.value.message
    with pytest.raises(sla.NeedsPivoting) as e:
        sla.triUpperSolve(sla.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 1, 1e-12), (2, 2, 1.0)]), one)
    assert "U (1,1)" in e.value.message
    # 1e-12 itself is nearZero (abs a <= 1e-12), the next double is not
    ok = sla.SpMatrix.fromListSM((3, 3), [(0, 0, 1.0), (1, 1, np.nextafter(1e-12, 1.0)), (2, 2, 1.0)])
    assert sla.triUpperSolve(ok, one).toDenseListSV()[1] == 1.0 / np.nextafter(1e-12, 1.0)
    # dimension 1: the reference's loop looks up (1,1) / (-1,-1) with the bounds-checked (@@)
    m1 = sla.SpMatrix.fromListSM((1, 1), [(0, 0, 2.0)])
    for f in (sla.triLowerSolve, sla.triUpperSolve):
        with pytest.raises(sla.OutOfBoundsIndexError):
            f(m1, sla.SpVector.mkSpVR(1, [4.0]))
    with pytest.raises(sla.NeedsPivoting):
        sla.triLowerSolve(sla.SpMatrix.fromListSM((1, 1), [(0, 0, 0.0)]), sla.SpVector.mkSpVR(1, [4.0]))
    with pytest.raises(sla.MatVecSizeMismatchException):
        sla.triLowerSolve(ok, sla.SpVector.mkSpVR(4, [1.0] * 4))
    with pytest.raises(sla.MatVecSizeMismatchException):
        sla.triLowerSolve(sla.SpMatrix.fromListSM((2, 3), [(0, 0, 1.0), (1, 1, 1.0)]), sla.SpVector.mkSpVR(2, [1.0, 1.0]))


# =============================================================== partitions and preconditioners, bit-exact

def csr_equal(A, Ao):
    rp, ci, va = A.toCSR()
    rpo, cio, vao = Ao.toCSR()
    return rp.tolist() == rpo.tolist() and ci.tolist() == cio.tolist() and va.tobytes() == vao.tobytes()


@pytest.mark.parametrize("kind,n,k,band", [("uniform", 2000, 9, 0), ("laplace", 30 * 30, 5, 30)])
def test_diag_partitions_and_jacobi(sla, o, kind, n, k, band):
    gk = {"uniform": (sla.GEN_UNIFORM, o.GEN_UNIFORM), "laplace": (sla.GEN_LAPLACE2D, o.GEN_LAPLACE2D)}[kind]
    A = sla.SpMatrix.generate(gk[0], n, k, 0x5EED0011, band)
    Ao = o.SpMatrix.synth(gk[1], n, k, 0x5EED0011, band)
    for got, want in zip(sla.diagPartitions(A), o.diagPartitions(Ao)):
        assert csr_equal(got, want)
    assert csr_equal(sla.jacobiPre(A), o.jacobiPre(Ao))
    assert csr_equal(A.extractSubDiag(), Ao.extractSubDiag()) and csr_equal(A.extractSuperDiag(), Ao.extractSuperDiag())
    # the parts are ordinary matrices: (e + d + f) #> x reproduces aa #> x up to the association of the row sums
    x = sla.SpVector.generate(n, 5)
    e, d, f = sla.diagPartitions(A)
    y = (e @ x) + (d @ x) + (f @ x)
    np.testing.assert_allclose(y.toDenseListSV(), (A @ x).toDenseListSV(), rtol=1e-12, atol=1e-12)


def test_partitions_rectangular_and_missing_diagonal(sla, o):
    rng = np.random.default_rng(73)
    for (m, n) in ((40, 25), (25, 40), (30, 30)):
        mask = rng.random((m, n)) < 0.2
        mask[np.arange(0, min(m, n), 3), np.arange(0, min(m, n), 3)] = False     # some diagonal entries missing
        i, j = np.nonzero(mask)
        v = rng.standard_normal(i.size)
        A, Ao = sla.SpMatrix.fromCOO((m, n), i, j, v), o.SpMatrix.fromCOO((m, n), i, j, v)
        for got, want in zip(sla.diagPartitions(A), o.diagPartitions(Ao)):
            assert csr_equal(got, want)
        assert csr_equal(sla.jacobiPre(A), o.jacobiPre(Ao))


@pytest.mark.parametrize("omega", [1.0, 1.5, 0.3])
def test_mssor_pre(sla, o, omega):
    """mSsorPre (Sparse.hs:713-721): values bit-identical; the reference's (##) additionally stores an explicit zero for
    every empty (row, column) intersection, which the CSR result does not materialise — compare as dense."""
    n, k = 180, 7
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, 0x5EED0012)
    Ao = o.SpMatrix.synth(o.GEN_UNIFORM, n, k, 0x5EED0012)
    l, r = sla.mSsorPre(A, omega)
    lo, ro = o.mSsorPre(Ao, omega)
    assert csr_equal(r, ro)
    assert same_bits(l.toDense(), lo.toDense())
    assert lo.nnz == n * n and l.nnz == A.extractSubDiag().nnz + n
    # with a missing diagonal entry the whole column disappears from l
    i, j, v = [0, 1, 2, 2, 3, 3], [0, 0, 1, 2, 1, 3], [2.0, 4.0, 3.0, 5.0, -1.0, 8.0]
    B, Bo = sla.SpMatrix.fromCOO((4, 4), i, j, v), o.SpMatrix.fromCOO((4, 4), i, j, v)
    l, r = sla.mSsorPre(B, omega)
    lo, ro = o.mSsorPre(Bo, omega)
    assert same_bits(l.toDense(), lo.toDense()) and same_bits(r.toDense(), ro.toDense())
    assert l.nnz == 4          # (0,0), (1,0), (2,2), (3,3): column 1 has no stored diagonal
    with pytest.raises(sla.MatVecSizeMismatchException):
        sla.mSsorPre(sla.SpMatrix.fromListSM((2, 3), [(0, 0, 1.0)]), 1.0)
    # symmetric Gauss-Seidel as a preconditioner application: z = r^-1 (l^-1 ... ) is NOT what the reference composes;
    # what it offers are the factors.  Applying them: solve r z = y by the backward sweep.
    y = sla.SpVector.generate(n, 9)
    l, r = sla.mSsorPre(A, omega)
    z = sla.triUpperSolve(r, y)
    np.testing.assert_allclose((r @ z).toDenseListSV(), y.toDenseListSV(), rtol=1e-10, atol=1e-12)


# =============================================================== BASELINE-size property (cfg3 matrix)

def test_full_size_laplacian_sweeps(sla):
    """n = 4096^2 five-point Laplacian: 8191 wavefronts.  Size-independent properties: (e + d) w = b and (d + f) x = b to
    rounding, and two runs give identical bits (the schedule is deterministic)."""
    side = 4096
    n = side * side
    A = sla.SpMatrix.generate(sla.GEN_LAPLACE2D, n, 5, 0x5EED0003, side)
    b = sla.SpVector.generate(n, 0x5EED0004)
    e, d, f = sla.diagPartitions(A)
    assert A.triAnalysis(False) == (2 * side - 1, n + 2 * side * (side - 1))
    w = sla.triLowerSolve(A, b)
    res = ((e @ w) + (d @ w)) - b
    assert res.norm2() <= 1e-12 * b.norm2()
    w2 = sla.triLowerSolve(A, b)
    assert (w - w2).norm2() == 0.0
    x = sla.triUpperSolve(A, b)
    res = ((d @ x) + (f @ x)) - b
    assert res.norm2() <= 1e-12 * b.norm2()


# ---- ilu0Pre (Sparse.hs:696-706): the reference's complete lu masked by aa's stored positions, bit-exact -------------------

@pytest.mark.parametrize("n,density,seed", [(1, 1.0, 0), (2, 1.0, 1), (5, 0.6, 2), (17, 0.4, 3), (50, 0.2, 4), (50, 1.0, 5), (300, 0.03, 6)])
def test_ilu0pre_bit_exact(sla, o, n, density, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) * (rng.random((n, n)) < density) + np.diag(rng.uniform(3, 5, n) * n ** 0.5)
    trip = [(i, j, A[i, j]) for i in range(n) for j in range(n) if A[i, j] != 0.0]
    lo, uo = o.ilu0Pre(o.SpMatrix.fromListSM((n, n), trip))
    lg, ug = sla.ilu0Pre(sla.SpMatrix.fromListSM((n, n), trip))
    assert csr_equal(lg, lo)
    assert csr_equal(ug, uo)


def test_ilu0pre_reference_fixtures_and_errors(sla, o):
    import math

    for dims, trip in (((2, 2), [(0, 0, 1.0), (1, 0, 3.0), (0, 1, 2.0), (1, 1, 4.0)]),
                       ((2, 2), [(0, 0, math.pi), (1, 0, math.sqrt(2)), (0, 1, math.e), (1, 1, math.sqrt(5))]), F.tm7_triples()):
        lo, uo = o.ilu0Pre(o.SpMatrix.fromListSM(dims, trip))
        lg, ug = sla.ilu0Pre(sla.SpMatrix.fromListSM(dims, trip))
        assert csr_equal(lg, lo) and csr_equal(ug, uo)
        # tridiagonal / dense 2 x 2: no fill-in, so the masked factors still multiply back to aa (checkLu, LibSpec.hs:424-434)
        d = lg.toDense() @ ug.toDense() - sla.SpMatrix.fromListSM(dims, trip).toDense()
        assert np.abs(d).max() <= 1e-12
    # explicit zeros stored in aa keep their positions in the mask; a stored 0 in column 0 is a stored 0 in L
    trip = [(0, 0, 2.0), (1, 0, 0.0), (1, 1, 3.0), (0, 1, 0.0), (2, 2, 1.0), (2, 0, 4.0)]
    lo, uo = o.ilu0Pre(o.SpMatrix.fromListSM((3, 3), trip))
    lg, ug = sla.ilu0Pre(sla.SpMatrix.fromListSM((3, 3), trip))
    assert csr_equal(lg, lo) and csr_equal(ug, uo)
    # NeedsPivoting: u00 = 0, and a pivot that cancels at step 1 (same pivot index as the oracle)
    for trip, dims in (([(0, 1, 1.0), (1, 0, 1.0)], (2, 2)),
                       ([(0, 0, 1.0), (0, 1, 1.0), (1, 0, 1.0), (1, 1, 1.0), (2, 2, 1.0), (2, 1, 1.0)], (3, 3))):
        with pytest.raises(o.NeedsPivoting) as eo:
            o.ilu0Pre(o.SpMatrix.fromListSM(dims, trip))
        with pytest.raises(sla.NeedsPivoting) as eg:
            sla.ilu0Pre(sla.SpMatrix.fromListSM(dims, trip))
        assert f"U({eo.value.row},{eo.value.row})" in str(eg.value)
    with pytest.raises(sla.SlaError):
        sla.ilu0Pre(sla.SpMatrix.generate(sla.GEN_UNIFORM, 5000, 4, 1))     # beyond the O(n^3) algorithm's range
