"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the
reference's own known-answer fixtures.  Run on the B200 box with `pytest -m gpu`.

Tolerances, u = 2^-53:
  * CSR construction / transpose / elementwise vector ops: BIT-EXACT.
  * SpMV rows of <= 256 entries: BIT-EXACT (products rounded once, summed in ascending column order from 0,
    exactly the reference's left fold).  Longer rows: |dy| <= (k + 2) u sum|a_ij x_j|.
  * dot / norm: |d| <= 64 u sum|x_i y_i| (tree reduction).
  * Krylov iterates: first steps within 1e-10 relative of the oracle trajectory; same iteration count.
"""
import ctypes as C

import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu
U = 2.0 ** -53


@pytest.fixture(scope="module")
def sla():
    import sparse_linear_algebra_b200 as s

    return s


@pytest.fixture(scope="module")
def o(ora):
    return ora


def dense(sla, fx):
    return sla.SpMatrix.fromListDenseSM(fx[0], fx[1])


def vr(sla, ll):
    return sla.SpVector.mkSpVR(len(ll), ll)


def nearZero(a):
    return abs(a) <= 1e-12


# =============================================================== the reference's own specs, on the device

def test_ref_dot(sla):                                  # LibSpec.hs:45-46
    tv0 = vr(sla, F.TV0)
    assert tv0.dot(tv0) == 61


def test_ref_transpose_exact(sla):                      # LibSpec.hs:49-50
    t = dense(sla, F.M1).transpose()
    e = dense(sla, F.M1T)
    for a, b in zip(t.toCSR(), e.toCSR()):
        assert a.tolist() == b.tolist()


def test_ref_matvec_vecmat(sla):                        # LibSpec.hs:51-54
    aa0, x0true = dense(sla, F.AA0), vr(sla, F.X0TRUE)
    assert (aa0 @ x0true).toDenseListSV().tolist() == [8.0, 18.0]
    assert aa0.vecMat(x0true).toDenseListSV().tolist() == [11.0, 16.0]
    assert nearZero(((aa0 @ x0true) - vr(sla, F.B0)).norm2Sq())


def test_ref_fromlist_semantics(sla):                   # LibSpec.hs:63-65, 1268 ; SpMatrix.hs:205-224
    m1p = sla.SpMatrix.fromListSM(*F.M1P)               # duplicate (1,2,4),(1,2,1): last write wins
    rp, ci, va = m1p.toCSR()
    assert rp.tolist() == [0, 1, 3] and ci.tolist() == [0, 0, 2] and va.tolist() == [2.0, 3.0, 1.0]
    with pytest.raises(sla.OutOfBoundsIndexError):
        sla.SpMatrix.fromListSM((2, 2), [(0, 2, 1.0)])
    with pytest.raises(sla.OutOfBoundsIndexError):
        sla.SpMatrix.fromListSM((2, 2), [(-1, 0, 1.0)])
    v = sla.SpVector.fromListSV(3, [(1, 7.0), (1, 9.0), (5, 1.0)])      # SpVector.hs:275-278
    assert v.toDenseListSV().tolist() == [0.0, 7.0, 0.0]


def test_ref_matvec_dim_mismatch(sla):                  # Common.hs:248-250
    with pytest.raises(sla.MatVecSizeMismatchException):
        dense(sla, F.AA0) @ vr(sla, [1, 2, 3])


def test_ref_eye_is_diagonal(sla):                      # LibSpec.hs:68-69 ; SpMatrix.hs:411-415
    e = sla.SpMatrix.eye(10)
    assert e.nnz == 10 and e.isDiagonalSM()
    assert not dense(sla, F.AA0).isDiagonalSM()
    # a stored row with a single OFF-diagonal entry is not diagonal
    assert not sla.SpMatrix.fromListSM((2, 2), [(0, 1, 1.0), (1, 0, 1.0)]).isDiagonalSM()


def test_ref_krylov_init_exact(sla, o):                 # LibSpec.hs:252-257, 265-269
    aa0, b0, x0 = dense(sla, F.AA0), vr(sla, F.B0), vr(sla, F.X0)
    r0 = (b0 - (aa0 @ x0)).toDenseListSV()
    st = sla.bicgsInit(aa0, b0, x0)
    assert st.r.toDenseListSV().tolist() == r0.tolist() and st.p.toDenseListSV().tolist() == r0.tolist()
    st = sla.cgsInit(aa0, b0, x0)
    for f in (st.r, st.p, st.u):
        assert f.toDenseListSV().tolist() == r0.tolist()
    # and identical to the oracle's b - A x0
    ao = o.SpMatrix.fromListDenseSM(*F.AA0)
    ro = o.SpVector.mkSpVR(2, F.B0) - ao.matVec(o.SpVector.mkSpVR(2, F.X0))
    assert ro.toDenseListSV().tolist() == r0.tolist()


def _check_solver(sla, init, step, aa, b, niter):
    """checkCGS / checkBiCGSTAB (LibSpec.hs:548-575, 606-632)."""
    x0 = sla.SpVector.fromListSV(b.dim, [])
    rhat = b - (aa @ x0)
    st = init(aa, b, x0)
    tol = max(1e-6, 1e-4 * st.r.norm2())
    res = lambda s: ((aa @ s.x) - b).norm2()
    n = 0
    while n < niter:
        st = step(aa, rhat, st)
        n += 1
        if res(st) <= tol:
            break
    return res(st) <= tol, n, st


@pytest.mark.parametrize("solver", ["cgs", "bicgstab"])
@pytest.mark.parametrize("system", ["aa0", "aa2"])
def test_ref_solver_converges(sla, solver, system):     # LibSpec.hs:259-262, 276-279
    aa = dense(sla, F.AA0 if system == "aa0" else F.AA2)
    b = vr(sla, F.B0 if system == "aa0" else F.B2)
    init, step = (sla.cgsInit, sla.cgsStep) if solver == "cgs" else (sla.bicgsInit, sla.bicgstabStep)
    ok, n, _ = _check_solver(sla, init, step, aa, b, 50)
    assert ok and n <= 50


@pytest.mark.parametrize("method", ["BICGSTAB_", "CGS_", "CGNE_"])
@pytest.mark.parametrize("system", ["aa0", "aa2"])
def test_ref_linsolve0(sla, o, method, system):         # LibSpec.hs:286-300: ||x - xhat|| <= 1e-12
    fx, bb, xx = (F.AA0, F.B0, F.X0TRUE) if system == "aa0" else (F.AA2, F.B2, F.X2)
    aa = dense(sla, fx)
    n = aa.ncols
    xhat, iters, _ = sla.linSolve0(getattr(sla, method), aa, vr(sla, bb), sla.SpVector.mkSpVR(n, [0.1] * n), info=True)
    assert nearZero((vr(sla, xx) - xhat).norm2())
    # same iteration count as the oracle
    ao = o.SpMatrix.fromListDenseSM(*fx)
    _, it_o, _ = o.linSolve0(getattr(o, method), ao, o.SpVector.mkSpVR(n, bb), o.SpVector.mkSpVR(n, [0.1] * n), info=True)
    assert iters == it_o


def test_ref_linsolve0_errors_and_diagonal(sla):        # Sparse.hs:1022, 1024-1025, 1031
    aa0, b0 = dense(sla, F.AA0), vr(sla, F.B0)
    x0r = sla.SpVector.mkSpVR(2, [0.1, 0.1])
    with pytest.raises(sla.IterE) as e:
        sla.linSolve0(sla.GMRES_, aa0, b0, x0r)
    assert "Only BICGSTAB_, CGS_, and CGNE_ are implemented, got: GMRES_" in str(e.value)
    with pytest.raises(sla.IterE):
        sla.linSolve0(sla.BCG_, aa0, b0, x0r)
    with pytest.raises(sla.MatVecSizeMismatchException):
        sla.linSolve0(sla.BICGSTAB_, aa0, vr(sla, [1, 2, 3]), x0r)
    d = sla.SpMatrix.fromListSM((3, 3), [(0, 0, 2.0), (1, 1, 4.0), (2, 2, 8.0)])
    x = sla.linSolve0(sla.BCG_, d, vr(sla, [2, 2, 2]), sla.SpVector.zeroSV(3))      # diagonal shortcut precedes the method check
    assert x.toDenseListSV().tolist() == [1.0, 0.5, 0.25]


def test_ref_readme_example(sla):                       # README.md:97, 183-241
    amat = sla.SpMatrix.fromListSM(*F.AMAT)
    b = vr(sla, F.AMAT_B)
    x = sla.linSolve0(sla.BICGSTAB_, amat, b, sla.SpVector.fromListSV(3, []))
    np.testing.assert_allclose(x.toDenseListSV(), F.AMAT_X, atol=1e-5)
    np.testing.assert_allclose((amat @ x).toDenseListSV(), F.AMAT_B, atol=1e-5)
    xg = sla.backslash(amat, b)                         # aa <\> b (GMRES)
    np.testing.assert_allclose(xg.toDenseListSV(), F.AMAT_X, atol=1e-5)


@pytest.mark.parametrize("which,kn", [("aa4", 3), ("tm7", 4)])
def test_ref_arnoldi(sla, o, which, kn):                # LibSpec.hs:226-232, checkArnoldi :638-653
    if which == "aa4":
        aa, ao = dense(sla, F.AA4), o.SpMatrix.fromListDenseSM(*F.AA4)
    else:
        aa, ao = sla.SpMatrix.fromListSM(*F.tm7_triples()), o.SpMatrix.fromListSM(*F.tm7_triples())
    n = aa.nrows
    Qd, H, brk = sla.arnoldi(aa, sla.SpVector.onesSV(n), kn)
    Q = Qd.toHost()
    A = aa.toDense()
    assert H.shape[0] == H.shape[1] + 1 and Q.shape == (n, H.shape[0])
    assert np.linalg.norm(A @ Q[:, :-1] - Q @ H) <= 1e-12
    Qo, Ho = o.arnoldi(ao, o.SpVector.onesSV(n), kn)
    assert Qo.shape == Q.shape and Ho.shape == H.shape
    # compare with the oracle where the process is well conditioned (before a breakdown column)
    good = H.shape[1] if not brk else H.shape[1] - 1
    np.testing.assert_allclose(H[:, :good], Ho[:, :good], atol=1e-9)


# =============================================================== oracle parity on seeded inputs

def _rand_coo(rng, m, n, k, long_rows=()):
    i = rng.integers(0, m, k)
    j = rng.integers(0, n, k)
    for r, ln in long_rows:
        i = np.concatenate([i, np.full(ln, r)])
        j = np.concatenate([j, rng.choice(n, ln, replace=False)])
    v = rng.standard_normal(i.size)
    return i, j, v


@pytest.mark.parametrize("seed,m,n,k", [(0, 1, 1, 1), (1, 7, 5, 0), (2, 50, 40, 300), (3, 3000, 2500, 20000),
                                        (4, 20000, 20000, 150000), (5, 100, 100000, 5000)])
def test_coo_to_csr_and_transpose_bit_exact(sla, o, seed, m, n, k):
    rng = np.random.default_rng(seed)
    i, j, v = _rand_coo(rng, m, n, k)
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    rp, ci, va = A.toCSR()
    rpo, cio, vao = Ao.toCSR()
    assert rp.tolist() == rpo.tolist() and ci.tolist() == cio.tolist()
    assert va.tobytes() == vao.tobytes()
    rp, ci, va = A.transpose().toCSR()
    rpo, cio, vao = Ao.transpose().toCSR()
    assert rp.tolist() == rpo.tolist() and ci.tolist() == cio.tolist()
    assert va.tobytes() == vao.tobytes()
    # transpose is an involution, bit for bit
    rp2, ci2, va2 = A.transpose().transpose().toCSR()
    rp1, ci1, va1 = A.toCSR()
    assert rp2.tolist() == rp1.tolist() and ci2.tolist() == ci1.tolist() and va2.tobytes() == va1.tobytes()


@pytest.mark.parametrize("seed,m,n,k", [(10, 1, 1, 1), (11, 9, 9, 0), (12, 64, 64, 500), (13, 5000, 4000, 40000),
                                        (14, 30000, 30000, 400000), (15, 300, 5000, 60000)])
def test_spmv_bit_exact_short_rows(sla, o, seed, m, n, k):
    """Rows of <= 256 entries: identical bits to the oracle's sequential left fold; includes empty rows, rows
    straddling tile boundaries (nnz > 2048) and matrices smaller than one tile."""
    rng = np.random.default_rng(seed)
    i, j, v = _rand_coo(rng, m, n, k)
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    rp = A.toCSR()[0]
    assert np.diff(rp).max(initial=0) <= 256
    x = rng.standard_normal(n)
    y = (A @ sla.SpVector.mkSpVR(n, x)).toDenseListSV()
    yo = Ao.matVec(o.SpVector.mkSpVR(n, x)).toDenseListSV()
    assert y.tobytes() == yo.tobytes()
    yh = A.matVecHost(x)                                 # host-buffer entry point, same bits
    assert yh.tobytes() == yo.tobytes()
    # (<#) through the cached transpose
    xt = rng.standard_normal(m)
    z = A.vecMat(sla.SpVector.mkSpVR(m, xt)).toDenseListSV()
    zo = Ao.vecMat(o.SpVector.mkSpVR(m, xt)).toDenseListSV()
    assert z.tobytes() == zo.tobytes()


def test_spmv_long_rows_within_bound(sla, o):
    """Rows longer than 256 entries (warp path) and longer than a tile: componentwise bound (k+2) u sum|a x|."""
    rng = np.random.default_rng(20)
    m, n = 400, 20000
    i, j, v = _rand_coo(rng, m, n, 3000, long_rows=[(3, 257), (77, 1000), (78, 2049), (200, 9000), (399, 5000)])
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    x = rng.standard_normal(n)
    y = (A @ sla.SpVector.mkSpVR(n, x)).toDenseListSV()
    yo = Ao.matVec(o.SpVector.mkSpVR(n, x)).toDenseListSV()
    rp, ci, va = A.toCSR()
    lens = np.diff(rp)
    absum = np.array([np.abs(va[rp[r]:rp[r + 1]] * x[ci[rp[r]:rp[r + 1]]]).sum() for r in range(m)])
    short = lens <= 256
    assert y[short].tobytes() == yo[short].tobytes()
    assert np.all(np.abs(y - yo) <= (lens + 2) * U * absum)
    assert (~short).sum() == 5


@pytest.mark.parametrize("panels", [2, 3, 7])
def test_spmv_column_panels_bit_exact(sla, o, monkeypatch, panels):
    """The column-panel layout (used when x outgrows L2) continues each row's left fold from panel to panel:
    short rows stay bit-identical, long rows within the bound, and the fused Krylov epilogues still work."""
    monkeypatch.setenv("SLA_SPMV_PANELS", str(panels))
    rng = np.random.default_rng(30 + panels)
    m, n = 6000, 5000
    i, j, v = _rand_coo(rng, m, n, 50000, long_rows=[(5, 300), (4000, 2500)])
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    x = rng.standard_normal(n)
    y = (A @ sla.SpVector.mkSpVR(n, x)).toDenseListSV()
    yo = Ao.matVec(o.SpVector.mkSpVR(n, x)).toDenseListSV()
    rp, ci, va = A.toCSR()
    lens = np.diff(rp)
    short = lens <= 256                      # every per-panel segment of such a row is summed sequentially
    assert y[short].tobytes() == yo[short].tobytes()
    assert A.matVecHost(x).tobytes() == y.tobytes()      # host-buffer entry: panel-wise upload, chunked download
    absum = np.array([np.abs(va[rp[r]:rp[r + 1]] * x[ci[rp[r]:rp[r + 1]]]).sum() for r in range(m)])
    assert np.all(np.abs(y - yo) <= (lens + 2) * U * absum)
    # Krylov epilogues (fused dots, residual norm) through the panel path: cfg-1 system, same iteration count
    nn, k, seed = 200, 9, 0x5EED0001
    S = sla.SpMatrix.generate(sla.GEN_UNIFORM, nn, k, seed)
    So = o.SpMatrix.synth(o.GEN_UNIFORM, nn, k, seed)
    bo = So.matVec(o.SpVector.synth(seed + 1, nn))
    xg, it, _ = sla.linSolve0(sla.BICGSTAB_, S, sla.SpVector.mkSpVR(nn, bo.toDenseListSV()), sla.SpVector.constv(nn, 0.1), info=True)
    xo, ito, _ = o.linSolve0(o.BICGSTAB_, So, bo, o.SpVector.mkSpVR(nn, [0.1] * nn), info=True)
    assert it == ito
    assert np.abs(xg.toDenseListSV() - xo.toDenseListSV()).max() <= 1e-10 * np.abs(xo.toDenseListSV()).max()


def test_synthetic_generators_match_oracle(sla, o):
    for kind, n, k, band in ((sla.GEN_UNIFORM, 3000, 32, 0), (sla.GEN_BANDED, 3000, 16, 100),
                             (sla.GEN_LAPLACE2D, 64 * 64, 5, 64), (sla.GEN_UNIFORM, 40, 64, 0)):
        A = sla.SpMatrix.generate(kind, n, k, 0x5EED0001, band)
        Ao = o.SpMatrix.synth(kind, n, k, 0x5EED0001, band)
        rp, ci, va = A.toCSR()
        rpo, cio, vao = Ao.toCSR()
        assert rp.tolist() == rpo.tolist() and ci.tolist() == cio.tolist() and va.tobytes() == vao.tobytes()
    x = sla.SpVector.generate(1000, 7).toDenseListSV()
    assert x.tobytes() == o.SpVector.synth(7, 1000).toDenseListSV().tobytes()


@pytest.mark.parametrize("n", [1, 2, 3, 1000, 1001, 1 << 20, (1 << 20) + 7])
def test_vector_ops(sla, o, n):
    rng = np.random.default_rng(n)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    va, vb = sla.SpVector.mkSpVR(n, a), sla.SpVector.mkSpVR(n, b)
    # elementwise: bit-exact
    assert (va + vb).toDenseListSV().tobytes() == (a + b).tobytes()
    assert (va - vb).toDenseListSV().tobytes() == (a - b).tobytes()
    assert (0.37 * va).toDenseListSV().tobytes() == (0.37 * a).tobytes()
    assert (va / 3.0).toDenseListSV().tobytes() == ((1.0 / 3.0) * a).tobytes()
    assert vb.axpy(-1.7, va).toDenseListSV().tobytes() == (b + (-1.7 * a)).tobytes()
    # reductions: tree order, toleranced against the oracle's sequential fold
    oa, ob = o.SpVector.mkSpVR(n, a), o.SpVector.mkSpVR(n, b)
    import math

    s = np.abs(a * b).sum()
    # hard bound against the oracle: both are n-term summations of the same products
    assert abs(va.dot(vb) - oa.dot(ob)) <= 2 * n * U * s + 1e-300
    assert abs(va.norm2Sq() - oa.norm2Sq()) <= 2 * n * U * (a * a).sum()
    assert abs(va.norm2() - oa.norm2()) <= 2 * n * U * oa.norm2()
    # tree-reduction quality against the exactly-summed products
    assert abs(va.dot(vb) - math.fsum((a * b).tolist())) <= 64 * U * s + 1e-300
    assert abs(va.norm2Sq() - math.fsum((a * a).tolist())) <= 64 * U * (a * a).sum()
    nv = va.normalize2().toDenseListSV()
    np.testing.assert_allclose(nv, oa.normalize2().toDenseListSV(), rtol=2 * n * U + 4 * U, atol=0)
    # deterministic: same bits on a second evaluation
    assert va.dot(vb) == va.dot(vb)


def _laplace_system(sla, o, g, seed):
    n = g * g
    A = sla.SpMatrix.generate(sla.GEN_LAPLACE2D, n, 5, 0, g)
    Ao = o.SpMatrix.synth(o.GEN_LAPLACE2D, n, 5, 0, g)
    xt = o.SpVector.synth(seed, n)
    bo = Ao.matVec(xt)
    b = sla.SpVector.mkSpVR(n, bo.toDenseListSV())
    return n, A, Ao, b, bo


@pytest.mark.parametrize("solver", ["bicgstab", "cgs", "cgne"])
def test_krylov_trajectory_laplace(sla, o, solver):
    """cfg 3 at reduced size (64^2 Laplacian): first 5 iterates within 1e-10 relative of the oracle."""
    n, A, Ao, b, bo = _laplace_system(sla, o, 64, 11)
    x0, x0o = sla.SpVector.zeroSV(n), o.SpVector.mkSpVR(n, np.zeros(n))
    if solver == "bicgstab":
        st, sto = sla.bicgsInit(A, b, x0), o.bicgsInit(Ao, bo, x0o)
    elif solver == "cgs":
        st, sto = sla.cgsInit(A, b, x0), o.cgsInit(Ao, bo, x0o)
    else:
        st, sto = sla.cgneInit(A, b, x0), o.cgneInit(Ao, bo, x0o)
    rhat, rhato = st.r.copy(), bo - Ao.matVec(x0o)
    assert st.r.toDenseListSV().tobytes() == sto.r.toDenseListSV().tobytes()
    for it in range(5):
        if solver == "bicgstab":
            sla.bicgstabStep(A, rhat, st); sto = o.bicgstabStep(Ao, rhato, sto)
        elif solver == "cgs":
            sla.cgsStep(A, rhat, st); sto = o.cgsStep(Ao, rhato, sto)
        else:
            sla.cgneStep(A, st); sto = o.cgneStep(Ao, sto)
        for f in ("x", "r", "p"):
            g_, o_ = getattr(st, f).toDenseListSV(), getattr(sto, f).toDenseListSV()
            assert np.abs(g_ - o_).max() <= 1e-10 * np.abs(o_).max(), (solver, it, f)


@pytest.mark.parametrize("method", ["BICGSTAB_", "CGS_", "CGNE_"])
def test_linsolve0_cfg1(sla, o, method):
    """BASELINE config 1: 200x200 diagonally dominant system, x0 = 0.1: same iteration count as the oracle and
    ||x_gpu - x_oracle||_inf <= 1e-10 ||x||_inf."""
    n, k, seed = 200, 9, 0x5EED0001
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    Ao = o.SpMatrix.synth(o.GEN_UNIFORM, n, k, seed)
    xt = o.SpVector.synth(seed + 1, n)
    bo = Ao.matVec(xt)
    b = sla.SpVector.mkSpVR(n, bo.toDenseListSV())
    x, it, res = sla.linSolve0(getattr(sla, method), A, b, sla.SpVector.constv(n, 0.1), info=True)
    xo, ito, hist = o.linSolve0(getattr(o, method), Ao, bo, o.SpVector.mkSpVR(n, [0.1] * n), info=True)
    assert it == ito
    xg, xr = x.toDenseListSV(), xo.toDenseListSV()
    assert np.abs(xg - xr).max() <= 1e-10 * np.abs(xr).max()
    assert abs(res - hist[-1]) <= 1e-8 * max(hist[-1], 1e-30) + 1e-16
    # host-buffer entry point gives the same answer
    xh = sla.linSolve0Host(getattr(sla, method), A, bo.toDenseListSV(), np.full(n, 0.1))
    assert xh.tobytes() == xg.tobytes()


def test_linsolve0_nan_runs_to_cap(sla):
    """No breakdown guards: a zero denominator yields NaN; NaN <= tol is False, so the loop runs
    max_iters and returns silently with SLA_OK."""
    aa = sla.SpMatrix.fromListSM((2, 2), [(0, 1, 1.0), (1, 0, -1.0)])        # r0hat . A r0hat = 0
    b = sla.SpVector.mkSpVR(2, [1.0, 1.0])
    x, it, res = sla.linSolve0(sla.BICGSTAB_, aa, b, sla.SpVector.zeroSV(2), nits=7, info=True)
    assert it == 7 and np.isnan(x.toDenseListSV()).all()


def test_arnoldi_cfg4_small(sla, o):
    """cfg 4 at reduced size: || A Q_k - Q_{k+1} H ||_F / ||A||_F <= 1e-12 sqrt(n), Q orthonormal, H vs oracle."""
    n, k, seed, kn = 2000, 16, 0x5EED0004, 30
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    Ao = o.SpMatrix.synth(o.GEN_UNIFORM, n, k, seed)
    b = sla.SpVector.generate(n, seed + 1)
    Qd, H, brk = sla.arnoldi(A, b, kn)
    assert not brk and H.shape == (kn + 1, kn) and Qd.dim == (n, kn + 1)
    Q = Qd.toHost()
    Ad = Ao.toDense()
    assert np.linalg.norm(Ad @ Q[:, :-1] - Q @ H) / np.linalg.norm(Ad) <= 1e-12 * np.sqrt(n)
    Qo, Ho = o.arnoldi(Ao, o.SpVector.synth(seed + 1, n), kn)
    # The reference orthogonalises with CLASSICAL Gram-Schmidt (Sparse.hs:655-657), which loses orthogonality
    # on this clustered-spectrum matrix (the oracle reaches |Q^T Q - I| ~ 0.7 by column 30).  Parity therefore
    # means: the same loss as the oracle, and agreement column by column while the basis is well conditioned.
    for j in (6, 10):
        eg = np.abs(Q[:, :j].T @ Q[:, :j] - np.eye(j)).max()
        eo = np.abs(Qo[:, :j].T @ Qo[:, :j] - np.eye(j)).max()
        assert eg <= 10 * eo + 1e-13
    np.testing.assert_allclose(H[:, :8], Ho[:, :8], atol=1e-9 * np.abs(Ho).max())
    np.testing.assert_allclose(Q[:, :8], Qo[:, :8], atol=1e-9)
    eg, eo = np.abs(Q.T @ Q - np.eye(kn + 1)).max(), np.abs(Qo.T @ Qo - np.eye(kn + 1)).max()
    assert 0.01 * eo <= eg <= 100 * eo


def test_gmres_converges(sla, o):
    n, k, seed = 5000, 16, 0x5EED0004
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    xt = sla.SpVector.generate(n, seed + 2)
    b = A @ xt
    x, it, res = sla.gmres(A, b, sla.SpVector.zeroSV(n), restart=30, tol_abs=1e-10, tol_rel=1e-12, info=True)
    assert res <= 1e-9 * b.norm2() + 1e-10
    assert ((A @ x) - b).norm2() <= 1e-9 * b.norm2() + 1e-10
    np.testing.assert_allclose(x.toDenseListSV(), xt.toDenseListSV(), atol=1e-8)


def test_gmres_restarts_and_solver_options(sla, o):
    """GMRES with a short restart length needs several cycles; linSolve0's check_every / recurrence-residual
    options change when the loop looks, not what it computes."""
    n, k, seed = 3000, 12, 0x5EED0007
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    xt = sla.SpVector.generate(n, seed + 2)
    b = A @ xt
    x, it, res = sla.gmres(A, b, sla.SpVector.zeroSV(n), restart=3, tol_abs=1e-10, tol_rel=1e-12, info=True)
    assert it > 3 and res <= 1e-9 * b.norm2() + 1e-10
    np.testing.assert_allclose(x.toDenseListSV(), xt.toDenseListSV(), atol=1e-8)
    x0 = sla.SpVector.constv(n, 0.1)
    xa, ita, _ = sla.linSolve0(sla.BICGSTAB_, A, b, x0, info=True)
    xb, itb, resb = sla.linSolve0(sla.BICGSTAB_, A, b, x0, check_every=3, info=True)
    assert itb >= ita and itb % 3 == 0 and itb - ita < 3
    xc, itc, resc = sla.linSolve0(sla.BICGSTAB_, A, b, x0, true_residual=False, info=True)
    assert abs(itc - ita) <= 1
    tol = max(1e-6, 1e-4 * (b - (A @ x0)).norm2())
    for xs in (xa, xb, xc):
        assert ((A @ xs) - b).norm2() <= 1.01 * tol


def test_csr_input_validation(sla, o):
    """sla_csr_from_csr: already-CSR input is accepted as is (bit-exact product) and validated."""
    import scipy.sparse as sp

    rng = np.random.default_rng(60)
    S = sp.random(300, 200, density=0.05, random_state=60, format="csr", dtype=np.float64)
    S.sort_indices()
    A = sla.SpMatrix.fromCSR(300, 200, S.indptr, S.indices, S.data)
    x = rng.standard_normal(200)
    Ao = o.SpMatrix.fromCSR(300, 200, S.indptr, S.indices, S.data)
    assert (A @ sla.SpVector.mkSpVR(200, x)).toDenseListSV().tobytes() == Ao.matVec(o.SpVector.mkSpVR(200, x)).toDenseListSV().tobytes()
    with pytest.raises(sla.SlaError):                    # columns not ascending
        sla.SpMatrix.fromCSR(1, 3, [0, 2], [2, 0], [1.0, 2.0])
    with pytest.raises(sla.OutOfBoundsIndexError):       # column out of range
        sla.SpMatrix.fromCSR(1, 3, [0, 1], [3], [1.0])
    with pytest.raises(sla.SlaError):                    # row_ptr does not end at nnz
        sla.SpMatrix.fromCSR(2, 3, [0, 1, 1], [0, 1], [1.0, 2.0])


# =============================================================== (##) with a dense right operand

def _bf16_round(a):
    """Round-to-nearest-even from float64 straight to bfloat16 (8 significant bits), returned as float64.
    (Going through float32 first would double-round about one value in 2^16.)"""
    m, e = np.frexp(np.asarray(a, dtype=np.float64))
    return np.ldexp(np.round(m * 256.0) / 256.0, e)


def test_ref_matmat_fixtures(sla):                       # LibSpec.hs:61-65, 1263-1271: exact ==
    m1 = dense(sla, F.M1)
    m2 = np.array([[5.0, 6.0], [7.0, 8.0]])
    assert (m1 @ sla.DenseMatrix.fromHost(m2)).toHost().tolist() == [[19.0, 22.0], [43.0, 50.0]]
    m1p = sla.SpMatrix.fromListSM(*F.M1P)                # [[2,0,0],[3,0,1]] after the duplicate overwrite
    m2p = np.array([[5.0, 3.0], [0.0, 0.0], [0.0, 4.0]])
    assert (m1p @ sla.DenseMatrix.fromHost(m2p)).toHost().tolist() == [[10.0, 6.0], [15.0, 13.0]]
    with pytest.raises(sla.MatVecSizeMismatchException):  # error "matMat : incompatible matrix sizes"  SpMatrix.hs:790-797
        m1 @ sla.DenseMatrix.fromHost(np.ones((3, 2)))


@pytest.mark.parametrize("seed,m,n,k,nnz", [(40, 1, 1, 1, 1), (41, 70, 50, 7, 600), (42, 2000, 1500, 33, 30000), (43, 300, 4000, 128, 20000)])
def test_spmm_f64_bit_exact(sla, o, seed, m, n, k, nnz):
    """C_ic = sum_k asc b_kc * a_ik with the products rounded once: identical bits to the oracle's matMat_."""
    rng = np.random.default_rng(seed)
    i, j, v = _rand_coo(rng, m, n, nnz, long_rows=[(0, min(n, 400))] if n >= 400 else ())
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    B = rng.standard_normal((n, k))
    C = (A @ sla.DenseMatrix.fromHost(B)).toHost()
    Bo = o.SpMatrix.fromListDenseSM(n, B.T.reshape(-1))   # column-major list, every entry stored
    Co = Ao.matMat(Bo).toDense()
    assert C.tobytes() == Co.tobytes()


def test_spmm_bf16_within_bound(sla, o):
    """cfg 5 at reduced size: bf16 A values and B, fp32 accumulation, bf16 C, against the fp64 oracle on the
    bf16-rounded inputs: |dC| <= 2^-8 |C| + (k_i + 2) 2^-24 sum|a||b|.  (bf16 keeps 8 significant bits, so the
    final rounding is 2^-8 relative; SURVEY.md §8d wrote 2^-9.)"""
    rng = np.random.default_rng(50)
    m = n = 4096
    k, nnz = 128, 32 * 4096
    i, j = rng.integers(0, m, nnz), rng.integers(0, n, nnz)
    v = _bf16_round(rng.uniform(-1, 1, nnz))
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    B = _bf16_round(rng.uniform(-1, 1, (n, k)))
    C = (A @ sla.DenseMatrix.fromHost(B, sla.BF16)).toHost()
    rp, ci, va = A.toCSR()
    import scipy.sparse as sp

    S = sp.csr_matrix((va, ci, rp), shape=(m, n))
    Cref = S @ B
    absum = abs(S) @ np.abs(B)
    lens = np.diff(rp)[:, None]
    bound = 2.0 ** -8 * np.abs(Cref) + (lens + 2) * 2.0 ** -24 * absum + 1e-30
    assert np.all(np.abs(C - Cref) <= bound)
    assert np.array_equal(C, _bf16_round(C))              # the output really is bf16


def _spmm_bound_check(A, Bh, C):
    import scipy.sparse as sp

    rp, ci, va = A.toCSR()
    S = sp.csr_matrix((va, ci, rp), shape=(A.nrows, A.ncols))
    Cref = S @ Bh
    absum = abs(S) @ np.abs(Bh)
    lens = np.diff(rp)[:, None]
    bound = 2.0 ** -8 * np.abs(Cref) + (lens + 2) * 2.0 ** -24 * absum + 1e-30
    assert np.all(np.abs(C - Cref) <= bound)
    assert np.array_equal(C, _bf16_round(C))


@pytest.mark.parametrize("pipe", ["1", "0"])
@pytest.mark.parametrize("m,n,nblocks,fill", [(16, 16, 1, 1.0), (64, 64, 6, 1.0), (160, 4096, 60, 0.6), (1000, 2000, 400, 1.0), (40000, 8192, 30000, 0.9)])
def test_spmm_tensor_core_path(sla, monkeypatch, m, n, nblocks, fill, pipe):
    """The tcgen05 / TMEM tile path (16 x 16 bf16 blocks, M128 N16 K16 MMAs, fp32 accumulators) on block-structured
    matrices, forced with SLA_SPMM_TC=1: same bound as the gather kernel; includes a ragged last block row, empty block
    rows and partially filled blocks (zero-padded).  pipe = 1: the TMA-fed warp-specialised pipeline (default);
    pipe = 0: the one-stage kernel it replaced."""
    monkeypatch.setenv("SLA_SPMM_TC", "1")
    monkeypatch.setenv("SLA_SPMM_TC_PIPE", pipe)
    rng = np.random.default_rng(m + n)
    nbr, nbc = (m + 15) // 16, n // 16
    ii, jj = [], []
    for _ in range(nblocks):
        br, bc = int(rng.integers(0, nbr)), int(rng.integers(0, nbc))
        r, c = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
        keep = (rng.random((16, 16)) < fill) & (br * 16 + r < m)
        ii.append((br * 16 + r)[keep]); jj.append((bc * 16 + c)[keep])
    i, j = np.concatenate(ii), np.concatenate(jj)
    v = _bf16_round(rng.uniform(-1, 1, i.size))
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Bh = _bf16_round(rng.uniform(-1, 1, (n, 128)))
    C = (A @ sla.DenseMatrix.fromHost(Bh, sla.BF16)).toHost()
    _spmm_bound_check(A, Bh, C)


def test_spmm_block16_family_both_paths(sla, monkeypatch):
    """cfg 5 'K16' family (two full 16 x 16 blocks per 16-row group): the plan picks the tensor-core path by
    itself; it and the gather kernel both stay within the bound, and agree to bf16 rounding."""
    n = 16 * 3000
    A = sla.SpMatrix.generate(sla.GEN_BLOCK16, n, 32, 0x5EED0005)
    assert A.nnz == 32 * n
    rng = np.random.default_rng(7)
    Bh = _bf16_round(rng.uniform(-1, 1, (n, 128)))
    Bd = sla.DenseMatrix.fromHost(Bh, sla.BF16)
    l0 = A.ctx.launches
    C_tc = (A @ Bd).toHost()
    monkeypatch.setenv("SLA_SPMM_TC", "0")
    C_g = (A @ Bd).toHost()
    # A's values are not bf16-representable here: compare against the product of the ROUNDED values
    rp, ci, va = A.toCSR()
    import scipy.sparse as sp

    S = sp.csr_matrix((_bf16_round(va), ci, rp), shape=(n, n))
    Cref = S @ Bh
    absum = abs(S) @ np.abs(Bh)
    bound = 2.0 ** -8 * np.abs(Cref) + 34 * 2.0 ** -24 * absum + 1e-30
    assert np.all(np.abs(C_tc - Cref) <= bound) and np.all(np.abs(C_g - Cref) <= bound)
    assert np.abs(C_tc - C_g).max() <= 2.0 ** -7 * np.abs(Cref).max()


# =============================================================== BASELINE sizes, size-independent properties

def _seq_row_dot(cols, vals, x):
    acc = 0.0
    for c, v in zip(cols.tolist(), vals.tolist()):
        acc = acc + v * x[c]
    return acc


@pytest.mark.parametrize("kind,band", [("uniform", 0), ("banded", 65536)])
def test_full_size_spmv_cfg2(sla, o, kind, band):
    """BASELINE config 2 (10M x 10M, 32 nnz/row): sampled rows regenerated by the oracle are bit-exact; linearity."""
    n, k, seed = 10_000_000, 32, 0x5EED0002
    gk = sla.GEN_UNIFORM if kind == "uniform" else sla.GEN_BANDED
    A = sla.SpMatrix.generate(gk, n, k, seed, band)
    assert A.nnz == n * k
    x = sla.SpVector.generate(n, seed + 1)
    y = (A @ x).toDenseListSV()
    xh = x.toDenseListSV()
    rng = np.random.default_rng(5)
    rows = np.concatenate([[0, 1, n - 1, n - 2, 63, 64, 65], rng.integers(0, n, 300)])
    for r in rows.tolist():
        cols, vals = o.synth_row(gk, n, k, seed, band, r)
        assert y[r] == _seq_row_dot(cols, vals, xh), r
    # linearity: A(2x) == 2 A x exactly (power-of-two scaling commutes with rounding)
    y2 = (A @ (2.0 * x)).toDenseListSV()
    assert y2.tobytes() == (2.0 * y).tobytes()
    # dot fused epilogue vs separate dot: one BiCGSTAB step keeps going without NaN
    assert np.isfinite(y).all()


def test_full_size_bicgstab_cfg3(sla, o):
    """BASELINE config 3 (5-point Laplacian 4096^2): A*1 is the boundary indicator; residual of b = A x_true drops."""
    g = 4096
    n = g * g
    A = sla.SpMatrix.generate(sla.GEN_LAPLACE2D, n, 5, 0, g)
    assert A.nnz == 5 * n - 4 * g
    y = (A @ sla.SpVector.onesSV(n)).toDenseListSV().reshape(g, g)
    exp = np.zeros((g, g))
    exp[0, :] += 1; exp[-1, :] += 1; exp[:, 0] += 1; exp[:, -1] += 1
    assert np.array_equal(y, exp)
    xt = sla.SpVector.generate(n, 3)
    b = A @ xt
    x0 = sla.SpVector.zeroSV(n)
    st = sla.bicgsInit(A, b, x0)
    rhat = st.r.copy()
    r0 = st.r.norm2()
    for _ in range(20):
        sla.bicgstabStep(A, rhat, st)
    # recurrence residual equals the true residual to rounding, and has decreased
    true_res = ((A @ st.x) - b).norm2()
    assert abs(true_res - st.r.norm2()) <= 1e-8 * r0
    assert true_res < 0.5 * r0


# =============================================================== round 2: pure Krylov records, fixed-work GMRES, Arnoldi bookkeeping on the device

def test_krylov_clone_gives_pure_steps(sla, o):
    """README.md:208 — `iterate (bicgstabStep amat r0hat) initState !! k` keeps initState alive: a pure step (clone, then
    advance the clone) leaves its argument untouched and follows the same trajectory as the in-place step."""
    n, k, seed = 3000, 12, 0x5EED0011
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    b = A @ sla.SpVector.generate(n, seed + 1)
    for init, step in ((sla.bicgsInit, sla.bicgstabStep), (sla.cgsInit, sla.cgsStep)):
        st0 = init(A, b, sla.SpVector.zeroSV(n))
        rhat = st0.r.copy()
        x0, r0, p0 = st0.x.toDenseListSV(), st0.r.toDenseListSV(), st0.p.toDenseListSV()
        seq = [st0]
        for _ in range(4):
            seq.append(step(A, rhat, seq[-1], pure=True))
        # the argument of every pure step is unchanged
        assert st0.x.toDenseListSV().tobytes() == x0.tobytes() and st0.r.toDenseListSV().tobytes() == r0.tobytes()
        assert st0.p.toDenseListSV().tobytes() == p0.tobytes()
        # same bits as the in-place recurrence
        st = init(A, b, sla.SpVector.zeroSV(n))
        for j in range(1, 5):
            step(A, rhat, st)
            assert st.x.toDenseListSV().tobytes() == seq[j].x.toDenseListSV().tobytes(), j
            assert st.r.toDenseListSV().tobytes() == seq[j].r.toDenseListSV().tobytes(), j
    stc = sla.cgneInit(A, b, sla.SpVector.zeroSV(n))
    st1 = sla.cgneStep(A, stc, pure=True)
    assert np.array_equal(stc.x.toDenseListSV(), np.zeros(n)) and np.abs(st1.x.toDenseListSV()).max() > 0


def test_rho_cache_survives_vector_reuse(sla):
    """the cached rho = r <.> r0hat must not be reused for a NEW shadow residual that the allocator happens to
    place where the freed one lived (stamps are unique per context now)."""
    n, k, seed = 2048, 8, 0x5EED0012
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    b = A @ sla.SpVector.generate(n, seed + 1)
    st = sla.bicgsInit(A, b, sla.SpVector.zeroSV(n))
    ref = sla.bicgsInit(A, b, sla.SpVector.zeroSV(n))
    rhat = st.r.copy()
    sla.bicgstabStep(A, rhat, st)
    sla.bicgstabStep(A, ref.r.copy(), ref)
    del rhat                                            # freed; the next vector of this size is likely to reuse the address
    other = 2.0 * ref.r.copy()                          # a different shadow residual
    sla.bicgstabStep(A, other, st)
    sla.bicgstabStep(A, 2.0 * ref.r.copy(), ref)        # reference: fresh vector, no cache hit possible by pointer
    assert st.x.toDenseListSV().tobytes() == ref.x.toDenseListSV().tobytes()


def test_gmres_fixed_work_runs_every_cycle(sla):
    """BASELINE config 4 is "GMRES(30), 10 restarts": with the stopping test off every cycle runs to its end."""
    n, k, seed = 4000, 16, 0x5EED0004
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    xt = sla.SpVector.generate(n, seed + 2)
    b = A @ xt
    x, it, res = sla.gmres(A, b, sla.SpVector.zeroSV(n), restart=10, nits=40, fixed_work=True, info=True)
    assert it == 40
    assert np.isfinite(res) and res <= 1e-6 * b.norm2()
    np.testing.assert_allclose(x.toDenseListSV(), xt.toDenseListSV(), atol=1e-6)
    # with the stopping test the same problem stops early, mid-cycle, with the same answer
    x2, it2, res2 = sla.gmres(A, b, sla.SpVector.zeroSV(n), restart=10, nits=40, tol_abs=1e-8, tol_rel=1e-12, info=True)
    assert it2 < 40 and res2 <= 1e-8
    np.testing.assert_allclose(x2.toDenseListSV(), xt.toDenseListSV(), atol=1e-6)


def test_arnoldi_long_basis_and_breakdown(sla, o):
    """More than 32 basis columns (two launches of the dot kernel per step) against the oracle, and a breakdown found after
    the whole run was queued: the columns computed past it are discarded, H is the reference's (nmax+1) x nmax block."""
    n, k, seed, kn = 1500, 10, 0x5EED0021, 40
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
    Ao = o.SpMatrix.synth(o.GEN_UNIFORM, n, k, seed)
    Qd, H, brk = sla.arnoldi(A, sla.SpVector.generate(n, seed + 1), kn)
    assert not brk and H.shape == (kn + 1, kn)
    Q = Qd.toHost()
    Ad = Ao.toDense()
    assert np.linalg.norm(Ad @ Q[:, :-1] - Q @ H) / np.linalg.norm(Ad) <= 1e-12 * np.sqrt(n)
    Qo, Ho = o.arnoldi(Ao, o.SpVector.synth(seed + 1, n), kn)
    np.testing.assert_allclose(H[:, :8], Ho[:, :8], atol=1e-9 * np.abs(Ho).max())
    # invariant subspace after 2 steps: diag(1, 2, 3, ...) applied to e0 + e1
    D = sla.SpMatrix.mkDiagonal(50, np.arange(1.0, 51.0))
    Do = o.SpMatrix.fromListSM((50, 50), [(i, i, float(i + 1)) for i in range(50)])
    v = np.zeros(50); v[0] = v[1] = 1.0
    Qb, Hb, brkb = sla.arnoldi(D, sla.SpVector.mkSpVR(50, v), 10)
    Qob, Hob = o.arnoldi(Do, o.SpVector.mkSpVR(50, v), 10)
    assert brkb and Hb.shape == Hob.shape and Qb.toHost().shape == Qob.shape
    good = Hb.shape[1] - 1
    np.testing.assert_allclose(Hb[:, :good], Hob[:, :good], atol=1e-12)


def test_spmv_bulk_staging_bit_exact(sla, o, monkeypatch):
    """The bulk-copy staged variant of the tile kernel (option "spmv_bulk": the (col, val) tile arrives through cp.async.bulk
    instead of LDG) computes the same bits: plain, column panels, long rows, fused Krylov epilogues."""
    ctx = sla.default_context()
    rng = np.random.default_rng(77)
    cases = []
    for (m, n, k) in ((1, 1, 1), (300, 200, 2500), (5000, 4000, 60000)):
        i, j, v = _rand_coo(rng, m, n, k, long_rows=((0, min(n, 700)),) if m > 1 else ())
        cases.append(((m, n), i, j, v))
    try:
        for dims, i, j, v in cases:
            x = rng.standard_normal(dims[1])
            for panels in ("1", "3"):
                monkeypatch.setenv("SLA_SPMV_PANELS", panels)
                A = sla.SpMatrix.fromCOO(dims, i, j, v)
                xs = sla.SpVector.mkSpVR(dims[1], x)
                ctx.set_option("spmv_bulk", 0)
                y0 = (A @ xs).toDenseListSV()
                ctx.set_option("spmv_bulk", 1)
                y1 = (A @ xs).toDenseListSV()
                assert y0.tobytes() == y1.tobytes(), (dims, panels)
        monkeypatch.delenv("SLA_SPMV_PANELS", raising=False)
        # a Krylov trajectory (fused dot epilogues ride on the same kernel)
        n, k, seed = 4096, 16, 0x5EED0031
        A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, seed)
        b = A @ sla.SpVector.generate(n, seed + 1)
        xs = []
        for bulk in (0, 1):
            ctx.set_option("spmv_bulk", bulk)
            st = sla.bicgsInit(A, b, sla.SpVector.zeroSV(n))
            rhat = st.r.copy()
            for _ in range(5):
                sla.bicgstabStep(A, rhat, st)
            xs.append(st.x.toDenseListSV())
        assert xs[0].tobytes() == xs[1].tobytes()
    finally:
        ctx.set_option("spmv_bulk", int(__import__("os").environ.get("SLA_SPMV_BULK", "0")))


# =============================================================== round 2: x staged through shared memory (band plan, spmv_band.cuh)

@pytest.mark.parametrize("form,R,W", [("1", "64", "32"), ("1", "16", "16"), ("1", "4096", "1024"), ("1", "8192", "4096"),
                                      ("3", "64", "32"), ("3", "32", "32"), ("3", "4096", "1024"), ("3", "8192", "8192")])
def test_spmv_band_plan_bit_exact(sla, o, monkeypatch, form, R, W):
    """The band plans (row blocks x column sub-panels, x in shared memory; SLA_SPMV_BAND=1: entry-sorted stream, spmv_band.cuh;
    =3: sliced-ELL cells with a thread per row, spmv_bandsell.cuh — both forced at any size here) give the same bits as the
    oracle's left fold: banded, stencil, a narrow random matrix with empty rows, long rows (any length is exact here), ragged
    sizes, odd n; with the Krylov epilogues riding on it (dots: tolerance, another reduction grid)."""
    monkeypatch.setenv("SLA_SPMV_BAND", form)
    monkeypatch.setenv("SLA_BAND_R", R)
    monkeypatch.setenv("SLA_BAND_W", W)
    seed = 0x5EED0051
    cases = [(sla.GEN_BANDED, o.GEN_BANDED, 30001, 16, 300), (sla.GEN_LAPLACE2D, o.GEN_LAPLACE2D, 97 * 97, 5, 97),
             (sla.GEN_BANDED, o.GEN_BANDED, 5000, 8, 7)]
    for gk, ok_, n, k, band in cases:
        A = sla.SpMatrix.generate(gk, n, k, seed, band)
        Ao = o.SpMatrix.synth(ok_, n, k, seed, band)
        x = sla.SpVector.generate(n, seed + 1)
        xo = o.SpVector.synth(seed + 1, n)
        assert (A @ x).toDenseListSV().tobytes() == Ao.matVec(xo).toDenseListSV().tobytes(), (gk, n, R, W)
    # hand-made: empty rows, one row of 900 entries inside a 1000-column window, a dense-ish cluster, odd dimensions
    rng = np.random.default_rng(11)
    m, n = 2501, 3001
    ii, jj = [], []
    for r in range(m):
        if r % 7 == 3:
            continue                                                     # empty row
        width = 900 if r == 1200 else int(rng.integers(1, 12))
        lo = max(0, min(n - 1000, r - 500))
        cols = rng.choice(1000, size=width, replace=False) + lo
        ii.append(np.full(width, r)); jj.append(cols)
    i, j = np.concatenate(ii), np.concatenate(jj)
    v = rng.standard_normal(i.size)
    A = sla.SpMatrix.fromCOO((m, n), i, j, v)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    xh = rng.standard_normal(n)
    y = (A @ sla.SpVector.mkSpVR(n, xh)).toDenseListSV()
    yo = Ao.matVec(o.SpVector.mkSpVR(n, xh)).toDenseListSV()
    assert y.tobytes() == yo.tobytes()
    # a BiCGSTAB trajectory on the banded family: same bits with and without the band plan
    n, k = 20000, 12
    xs = []
    for bandplan in (form, "0"):
        monkeypatch.setenv("SLA_SPMV_BAND", bandplan)
        A = sla.SpMatrix.generate(sla.GEN_BANDED, n, k, seed, 200)
        b = A @ sla.SpVector.generate(n, seed + 2)
        st = sla.bicgsInit(A, b, sla.SpVector.zeroSV(n))
        rhat = st.r.copy()
        for _ in range(5):
            sla.bicgstabStep(A, rhat, st)
        xs.append(st.x.toDenseListSV())
        xsol, its, res = sla.linSolve0(sla.BICGSTAB_, A, b, sla.SpVector.constv(n, 0.1), info=True)
        xs.append(np.array([its, res]))
    # (#>) is bit-identical either way; the fused dots are reduced over a different grid, so the iterates agree to rounding
    np.testing.assert_allclose(xs[0], xs[2], rtol=1e-10, atol=1e-13)
    assert xs[1][0] == xs[3][0] and abs(xs[1][1] - xs[3][1]) <= 1e-8 * max(abs(xs[3][1]), 1e-300)


# =============================================================== round 2: rotated column panels of the phased x exchange (p2p.cu mode 5)
@pytest.mark.parametrize("world,spec", [(2, None), (4, None), (4, "1,3"), (8, None), (8, "1,1,1,1,2,2"), (5, "1,2,2")])
def test_spmv_rotated_panels_single_gpu(sla, o, world, spec):
    """The panel builder behind SLA_P2P_X=5, exercised on ONE GPU through the test hook: the matrix is cut as if its columns were
    `world` equal blocks and this GPU held block `rank`; (#>) then folds each row panel by panel, own block first.  Every entry
    must land in exactly one panel (checked through the result: each row within (k + 2) u sum |a_ij x_j| of the oracle's left
    fold, rows living in a single panel bit-exact), for every rank's rotation, wrap-around included."""
    ctx = sla.default_context()
    rng = np.random.default_rng(1000 + world)
    n = 40 * world                                           # 40 columns per block
    m = 333
    i, j, v = _rand_coo(rng, m, n, 9000, long_rows=((0, min(n, 200)),))      # every row <= 256 entries: the one-pass product is bit-exact
    # rows 1..10 live entirely inside one column block each: one panel only, so they stay bit-exact under any rotation
    keep = ~((i >= 1) & (i <= 10))
    i, j, v = i[keep], j[keep], v[keep]
    for r in range(1, 11):
        blk = (r * 3) % world
        cols = np.sort(rng.choice(40, size=17, replace=False)) + 40 * blk
        i = np.concatenate([i, np.full(17, r)]); j = np.concatenate([j, cols]); v = np.concatenate([v, rng.standard_normal(17)])
    x = rng.standard_normal(n)
    Ao = o.SpMatrix.fromCOO((m, n), i, j, v)
    yo = Ao.matVec(o.SpVector.mkSpVR(n, x)).toDenseListSV()
    rp, cj, vv = Ao.toCSR()
    mag = np.add.reduceat(np.abs(vv) * np.abs(x)[cj], rp[:-1])
    mag[np.diff(rp) == 0] = 0.0
    bound = (np.diff(rp) + 2) * 2.0 ** -53 * mag
    sizes = (C.c_int * 8)()
    npan = ctx.lib.sla_p2p_phase_schedule(world, spec.encode() if spec else None, sizes)
    xs = sla.SpVector.mkSpVR(n, x)
    for rank in range(world):
        A = sla.SpMatrix.fromCOO((m, n), i, j, v)
        base = (A @ xs).toDenseListSV()
        assert base.tobytes() == yo.tobytes()
        ctx.check(ctx.lib.sla_csr_debug_rot_panels(ctx.h, A.h, world, rank, spec.encode() if spec else None))
        assert ctx.lib.sla_csr_npanels(A.h) == npan
        y = (A @ xs).toDenseListSV()
        assert np.all(np.abs(y - yo) <= bound), (world, spec, rank, float(np.max(np.abs(y - yo) / np.maximum(bound, 1e-300))))
        assert y[1:11].tobytes() == yo[1:11].tobytes(), (world, spec, rank)
        ctx.check(ctx.lib.sla_csr_debug_rot_panels(ctx.h, A.h, 1, 0, None))       # panels off again: one pass, bit-exact
        assert ctx.lib.sla_csr_npanels(A.h) == 0
        assert (A @ xs).toDenseListSV().tobytes() == yo.tobytes()
    # a Krylov trajectory on rotated panels (the fused dot epilogues ride on the LAST panel's launch)
    nk = 4096
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, nk, 16, 0x5EED0041)
    b = A @ sla.SpVector.generate(nk, 0x5EED0042)
    traj = []
    for rot in (False, True):
        if rot:
            ctx.check(ctx.lib.sla_csr_debug_rot_panels(ctx.h, A.h, world if nk % world == 0 else 4, 1, None))
        st = sla.bicgsInit(A, b, sla.SpVector.zeroSV(nk))
        rhat = st.r.copy()
        for _ in range(5):
            sla.bicgstabStep(A, rhat, st)
        traj.append(st.x.toDenseListSV())
    assert np.allclose(traj[0], traj[1], rtol=1e-9, atol=1e-12)
