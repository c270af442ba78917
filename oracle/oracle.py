"""ctypes wrapper over the CPU oracle (oracle/libsla_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Names follow the reference
(ocramz/sparse-linear-algebra): SpVector / SpMatrix, matVec (#>), vecMat (<#), dot (<.>),
bicgsInit / bicgstabStep, cgsInit / cgsStep, cgneInit / cgneStep, arnoldi, linSolve0.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsla_oracle.so")

ORA_OK, ORA_ERR_SIZE_MISMATCH, ORA_ERR_OOB_INDEX, ORA_ERR_UNSUPPORTED_METHOD, ORA_ERR_NEEDS_PIVOTING = 0, 1, 2, 3, 4
GMRES_, CGNE_, BCG_, CGS_, BICGSTAB_ = 0, 1, 2, 3, 4
GEN_UNIFORM, GEN_BANDED, GEN_LAPLACE2D, GEN_BLOCK16 = 0, 1, 2, 3


class OracleError(Exception):
    def __init__(self, code, what):
        super().__init__(f"{what}: oracle error {code}")
        self.code = code


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("sla_oracle.c", "sla_oracle.h", "../include/sla_synth.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_p = C.c_void_p
_i64 = C.c_int64
_f64 = C.c_double
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)
_pint = C.POINTER(C.c_int)


class _Krylov(C.Structure):
    _fields_ = [("x", _p), ("r", _p), ("p", _p), ("u", _p)]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO)
    sig = {
        "ora_sv_zero": (_p, [_i64]),
        "ora_sv_from_dense": (_p, [_i64, _pf64, _i64]),
        "ora_sv_from_list": (_p, [_i64, _i64, _pi64, _pf64]),
        "ora_sv_copy": (_p, [_p]),
        "ora_sv_free": (None, [_p]),
        "ora_sv_dim": (_i64, [_p]),
        "ora_sv_nnz": (_i64, [_p]),
        "ora_sv_to_dense": (None, [_p, _pf64]),
        "ora_sv_to_list": (None, [_p, _pi64, _pf64]),
        "ora_sv_add": (_p, [_p, _p]),
        "ora_sv_negate": (_p, [_p]),
        "ora_sv_sub": (_p, [_p, _p]),
        "ora_sv_scale": (_p, [_f64, _p]),
        "ora_sv_divs": (_p, [_p, _f64]),
        "ora_sv_dot": (_f64, [_p, _p]),
        "ora_sv_norm2sq": (_f64, [_p]),
        "ora_sv_norm2": (_f64, [_p]),
        "ora_sv_normalize2": (_p, [_p]),
        "ora_near_zero": (C.c_int, [_f64]),
        "ora_sm_zero": (_p, [_i64, _i64]),
        "ora_sm_from_list": (_p, [_i64, _i64, _i64, _pi64, _pi64, _pf64, _pint]),
        "ora_sm_from_dense_colmajor": (_p, [_i64, _pf64, _i64]),
        "ora_sm_from_csr": (_p, [_i64, _i64, _pi64, _pi64, _pf64]),
        "ora_sm_free": (None, [_p]),
        "ora_sm_nrows": (_i64, [_p]),
        "ora_sm_ncols": (_i64, [_p]),
        "ora_sm_nnz": (_i64, [_p]),
        "ora_sm_nstored_rows": (_i64, [_p]),
        "ora_sm_to_coo": (None, [_p, _pi64, _pi64, _pf64]),
        "ora_sm_to_csr": (None, [_p, _pi64, _pi64, _pf64]),
        "ora_sm_transpose": (_p, [_p]),
        "ora_sm_is_diagonal": (C.c_int, [_p]),
        "ora_sm_reciprocal": (_p, [_p]),
        "ora_sm_sparsify": (_p, [_p]),
        "ora_sm_matvec": (_p, [_p, _p, _pint]),
        "ora_sm_vecmat": (_p, [_p, _p, _pint]),
        "ora_sm_matmat": (_p, [_p, _p, _pint]),
        "ora_sm_equal": (C.c_int, [_p, _p]),
        "ora_sm_extract_tri": (_p, [_p, C.c_int]),
        "ora_sm_eye": (_p, [_i64]),
        "ora_sm_scale_right": (_p, [_p, _f64]),
        "ora_sm_negate": (_p, [_p]),
        "ora_sm_add": (_p, [_p, _p]),
        "ora_sm_sub": (_p, [_p, _p]),
        "ora_jacobi_pre": (_p, [_p]),
        "ora_mssor_pre": (C.c_int, [_p, _f64, C.POINTER(_p), C.POINTER(_p)]),
        "ora_lu": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(C.c_int64)]),
        "ora_ilu0_pre": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(C.c_int64)]),
        "ora_tri_lower_solve": (_p, [_p, _p, _pint, _pi64]),
        "ora_tri_upper_solve": (_p, [_p, _p, _pint, _pi64]),
        "ora_krylov_free": (None, [C.POINTER(_Krylov)]),
        "ora_bicgs_init": (C.POINTER(_Krylov), [_p, _p, _p]),
        "ora_bicgstab_step": (C.POINTER(_Krylov), [_p, _p, C.POINTER(_Krylov)]),
        "ora_cgs_init": (C.POINTER(_Krylov), [_p, _p, _p]),
        "ora_cgs_step": (C.POINTER(_Krylov), [_p, _p, C.POINTER(_Krylov)]),
        "ora_cgne_init": (C.POINTER(_Krylov), [_p, _p, _p]),
        "ora_cgne_step": (C.POINTER(_Krylov), [_p, C.POINTER(_Krylov)]),
        "ora_linsolve0": (_p, [C.c_int, _p, _p, _p, C.c_int, _f64, _f64, _pint, _pf64, _pint]),
        "ora_arnoldi": (C.c_int, [_p, _p, C.c_int, C.c_int, _pf64, _pf64, _pint, _pint]),
        "ora_synth_matrix": (_p, [C.c_int, _i64, C.c_int, C.c_uint64, _i64]),
        "ora_synth_vector": (_p, [C.c_uint64, _i64]),
        "ora_synth_row": (None, [C.c_int, _i64, C.c_int, C.c_uint64, _i64, _i64, _pi64, _pf64, _pint]),
        "ora_time_matvec": (_f64, [_p, _p, C.c_int, C.c_int, _pf64]),
        "ora_time_bicgstab": (_f64, [_p, _p, _p, C.c_int, _pf64]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _f64arr(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_pf64)


def _i64arr(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_pi64)


class SpVector:
    """SpVector Double (src/Data/Sparse/SpVector.hs:42-43)."""

    def __init__(self, handle):
        assert handle, "null oracle vector"
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ora_sv_free(self._h)
            self._h = None

    # -- construction
    @staticmethod
    def zeroSV(n):
        return SpVector(lib().ora_sv_zero(n))

    @staticmethod
    def mkSpVR(d, ll):
        a, p = _f64arr(ll)
        return SpVector(lib().ora_sv_from_dense(d, p, a.size))

    fromListDenseSV = mkSpVR

    @staticmethod
    def fromListSV(d, iix):
        iix = list(iix)
        ia, ip = _i64arr([t[0] for t in iix])
        va, vp = _f64arr([t[1] for t in iix])
        return SpVector(lib().ora_sv_from_list(d, len(iix), ip, vp))

    @staticmethod
    def onesSV(d):
        return SpVector.mkSpVR(d, np.ones(d))

    @staticmethod
    def synth(seed, n):
        return SpVector(lib().ora_synth_vector(seed, n))

    # -- inspection
    @property
    def dim(self):
        return lib().ora_sv_dim(self._h)

    @property
    def nnz(self):
        return lib().ora_sv_nnz(self._h)

    def toDenseListSV(self):
        out = np.zeros(self.dim, dtype=np.float64)
        lib().ora_sv_to_dense(self._h, out.ctypes.data_as(_pf64))
        return out

    def toListSV(self):
        n = self.nnz
        idx = np.zeros(n, dtype=np.int64)
        val = np.zeros(n, dtype=np.float64)
        lib().ora_sv_to_list(self._h, idx.ctypes.data_as(_pi64), val.ctypes.data_as(_pf64))
        return list(zip(idx.tolist(), val.tolist()))

    # -- algebra (Class.hs:57-99)
    def __add__(self, w):      # ^+^
        return SpVector(lib().ora_sv_add(self._h, w._h))

    def __sub__(self, w):      # ^-^
        return SpVector(lib().ora_sv_sub(self._h, w._h))

    def __neg__(self):         # negateV
        return SpVector(lib().ora_sv_negate(self._h))

    def __rmul__(self, a):     # a .* v
        return SpVector(lib().ora_sv_scale(float(a), self._h))

    def __truediv__(self, s):  # v ./ s
        return SpVector(lib().ora_sv_divs(self._h, float(s)))

    def dot(self, w):          # <.>
        return lib().ora_sv_dot(self._h, w._h)

    def norm2Sq(self):
        return lib().ora_sv_norm2sq(self._h)

    def norm2(self):
        return lib().ora_sv_norm2(self._h)

    def normalize2(self):
        return SpVector(lib().ora_sv_normalize2(self._h))

    def __eq__(self, w):       # derived Eq
        return self.dim == w.dim and self.toListSV() == w.toListSV()


def nearZero(a):
    return bool(lib().ora_near_zero(float(a)))


class SpMatrix:
    """SpMatrix Double (src/Data/Sparse/SpMatrix.hs:52-54)."""

    def __init__(self, handle):
        assert handle, "null oracle matrix"
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ora_sm_free(self._h)
            self._h = None

    @staticmethod
    def fromListSM(dims, iix):
        iix = list(iix)
        ia, ip = _i64arr([t[0] for t in iix])
        ja, jp = _i64arr([t[1] for t in iix])
        va, vp = _f64arr([t[2] for t in iix])
        err = C.c_int(0)
        h = lib().ora_sm_from_list(dims[0], dims[1], len(iix), ip, jp, vp, C.byref(err))
        if err.value:
            raise OracleError(err.value, "insertSpMatrix : index out of bounds")
        return SpMatrix(h)

    @staticmethod
    def fromCOO(dims, i, j, v):
        ia, ip = _i64arr(i)
        ja, jp = _i64arr(j)
        va, vp = _f64arr(v)
        err = C.c_int(0)
        h = lib().ora_sm_from_list(dims[0], dims[1], ia.size, ip, jp, vp, C.byref(err))
        if err.value:
            raise OracleError(err.value, "insertSpMatrix : index out of bounds")
        return SpMatrix(h)

    @staticmethod
    def fromListDenseSM(m, ll):
        a, p = _f64arr(ll)
        return SpMatrix(lib().ora_sm_from_dense_colmajor(m, p, a.size))

    @staticmethod
    def fromCSR(m, n, row_ptr, col, val):
        ra, rp = _i64arr(row_ptr)
        ca, cp = _i64arr(col)
        va, vp = _f64arr(val)
        return SpMatrix(lib().ora_sm_from_csr(m, n, rp, cp, vp))

    @staticmethod
    def eye(n):
        return SpMatrix.fromListSM((n, n), [(i, i, 1.0) for i in range(n)])

    @staticmethod
    def mkSubDiagonal(n, o, xx):
        ii = list(range(n))
        jj = list(range(abs(o), n))
        a, b = (ii, jj) if o >= 0 else (jj, ii)
        return SpMatrix.fromListSM((n, n), list(zip(a, b, xx)))

    @staticmethod
    def synth(kind, n, k, seed, band=0):
        return SpMatrix(lib().ora_synth_matrix(kind, n, k, seed, band))

    @property
    def nrows(self):
        return lib().ora_sm_nrows(self._h)

    @property
    def ncols(self):
        return lib().ora_sm_ncols(self._h)

    @property
    def dim(self):
        return (self.nrows, self.ncols)

    @property
    def nnz(self):
        return lib().ora_sm_nnz(self._h)

    def toCOO(self):
        n = self.nnz
        i = np.zeros(n, dtype=np.int64)
        j = np.zeros(n, dtype=np.int64)
        v = np.zeros(n, dtype=np.float64)
        lib().ora_sm_to_coo(self._h, i.ctypes.data_as(_pi64), j.ctypes.data_as(_pi64), v.ctypes.data_as(_pf64))
        return i, j, v

    def toCSR(self):
        n = self.nnz
        rp = np.zeros(self.nrows + 1, dtype=np.int64)
        c = np.zeros(n, dtype=np.int64)
        v = np.zeros(n, dtype=np.float64)
        lib().ora_sm_to_csr(self._h, rp.ctypes.data_as(_pi64), c.ctypes.data_as(_pi64), v.ctypes.data_as(_pf64))
        return rp, c, v

    def toDense(self):
        d = np.zeros(self.dim)
        i, j, v = self.toCOO()
        d[i, j] = v
        return d

    def transpose(self):
        return SpMatrix(lib().ora_sm_transpose(self._h))

    def isDiagonalSM(self):
        return bool(lib().ora_sm_is_diagonal(self._h))

    def reciprocal(self):
        return SpMatrix(lib().ora_sm_reciprocal(self._h))

    def sparsifySM(self):
        return SpMatrix(lib().ora_sm_sparsify(self._h))

    def matVec(self, x):       # aa #> x
        err = C.c_int(0)
        h = lib().ora_sm_matvec(self._h, x._h, C.byref(err))
        if err.value:
            raise OracleError(err.value, "matVec : mismatched dimensions")
        return SpVector(h)

    def vecMat(self, x):       # x <# aa
        err = C.c_int(0)
        h = lib().ora_sm_vecmat(x._h, self._h, C.byref(err))
        if err.value:
            raise OracleError(err.value, "vecMat : mismatching dimensions")
        return SpVector(h)

    def matMat(self, b):       # aa ## b
        err = C.c_int(0)
        h = lib().ora_sm_matmat(self._h, b._h, C.byref(err))
        if err.value:
            raise OracleError(err.value, "matMat : incompatible matrix sizes")
        return SpMatrix(h)

    def __matmul__(self, o):
        return self.matVec(o) if isinstance(o, SpVector) else self.matMat(o)

    def __eq__(self, b):
        return bool(lib().ora_sm_equal(self._h, b._h))

    # -- diagonal partitions and matrix arithmetic used by the preconditioners (Sparse.hs:670-721)
    def extractSubDiag(self):
        return SpMatrix(lib().ora_sm_extract_tri(self._h, -1))

    def extractDiag(self):
        return SpMatrix(lib().ora_sm_extract_tri(self._h, 0))

    def extractSuperDiag(self):
        return SpMatrix(lib().ora_sm_extract_tri(self._h, 1))

    def scale(self, n):        # scale n = fmap (* n)
        return SpMatrix(lib().ora_sm_scale_right(self._h, float(n)))

    def __add__(self, b):
        return SpMatrix(lib().ora_sm_add(self._h, b._h))

    def __sub__(self, b):
        return SpMatrix(lib().ora_sm_sub(self._h, b._h))

    def __neg__(self):
        return SpMatrix(lib().ora_sm_negate(self._h))


class KrylovState:
    """BICGSTAB / CGS / CGNE record (Sparse.hs:855-963). Fields are oracle SpVectors."""

    def __init__(self, ptr):
        assert ptr, "null krylov state"
        self._ptr = ptr

    def __del__(self):
        if getattr(self, "_ptr", None) and _lib is not None:
            _lib.ora_krylov_free(self._ptr)
            self._ptr = None

    def _field(self, name):
        h = getattr(self._ptr.contents, name)
        return SpVector(lib().ora_sv_copy(h)) if h else None

    x = property(lambda s: s._field("x"))
    r = property(lambda s: s._field("r"))
    p = property(lambda s: s._field("p"))
    u = property(lambda s: s._field("u"))


class NeedsPivoting(OracleError):
    """MatrixException NeedsPivoting (Control/Exception/Common.hs:57-61)."""

    def __init__(self, what, row):
        super().__init__(ORA_ERR_NEEDS_PIVOTING, f"{what} : ({row},{row}) is close to 0")
        self.row = row


def diagPartitions(aa):        # Sparse.hs:673-679
    return aa.extractSubDiag(), aa.extractDiag(), aa.extractSuperDiag()


def jacobiPre(aa):             # Sparse.hs:686-687
    return SpMatrix(lib().ora_jacobi_pre(aa._h))


def mSsorPre(aa, omega):       # Sparse.hs:713-721
    l, r = _p(), _p()
    err = lib().ora_mssor_pre(aa._h, float(omega), C.byref(l), C.byref(r))
    if err:
        raise OracleError(err, "matMat : incompatible matrix sizes")
    return SpMatrix(l.value), SpMatrix(r.value)


def _lu(fn, what, aa):
    l, u, bad = _p(), _p(), C.c_int64(-1)
    err = fn(aa._h, C.byref(l), C.byref(u), C.byref(bad))
    if err == 4:
        raise NeedsPivoting(what, bad.value)
    if err:
        raise OracleError(err, what)
    return SpMatrix(l.value), SpMatrix(u.value)


def lu(aa):                    # Sparse.hs:489-538
    return _lu(lib().ora_lu, "solveForLij", aa)


def ilu0Pre(aa):               # Sparse.hs:696-706
    return _lu(lib().ora_ilu0_pre, "solveForLij", aa)


def _tri(fn, what, mm, v):
    err, bad = C.c_int(0), C.c_int64(-1)
    h = fn(mm._h, v._h, C.byref(err), C.byref(bad))
    if err.value == ORA_ERR_NEEDS_PIVOTING:
        raise NeedsPivoting(what, bad.value)
    if err.value:
        raise OracleError(err.value, "@@ : incompatible indices")
    return SpVector(h)


def triLowerSolve(ll, b):      # Sparse.hs:750-777
    return _tri(lib().ora_tri_lower_solve, "triLowerSolve", ll, b)


def triUpperSolve(uu, w):      # Sparse.hs:784-811
    return _tri(lib().ora_tri_upper_solve, "triUpperSolve", uu, w)


def bicgsInit(aa, b, x0):
    return KrylovState(lib().ora_bicgs_init(aa._h, b._h, x0._h))


def bicgstabStep(aa, r0hat, st):
    return KrylovState(lib().ora_bicgstab_step(aa._h, r0hat._h, st._ptr))


def cgsInit(aa, b, x0):
    return KrylovState(lib().ora_cgs_init(aa._h, b._h, x0._h))


def cgsStep(aa, rhat, st):
    return KrylovState(lib().ora_cgs_step(aa._h, rhat._h, st._ptr))


def cgneInit(aa, b, x0):
    return KrylovState(lib().ora_cgne_init(aa._h, b._h, x0._h))


def cgneStep(aa, st):
    return KrylovState(lib().ora_cgne_step(aa._h, st._ptr))


def linSolve0(method, aa, b, x0, nits=0, tol_abs=0.0, tol_rel=0.0, info=False):
    iters = C.c_int(0)
    err = C.c_int(0)
    hist = np.zeros(nits if nits > 0 else 200, dtype=np.float64)
    h = lib().ora_linsolve0(method, aa._h, b._h, x0._h, nits, tol_abs, tol_rel, C.byref(iters),
                            hist.ctypes.data_as(_pf64), C.byref(err))
    if err.value:
        raise OracleError(err.value, "linSolve0")
    x = SpVector(h)
    return (x, iters.value, hist[: iters.value].copy()) if info else x


def arnoldi(aa, b, kn, max_steps=None):
    """Returns (Q dense n x ncols, H dense (nmax+1) x nmax), both numpy, as `arnoldi` (Sparse.hs:630-667)."""
    n = aa.nrows
    if max_steps is None:
        max_steps = kn - 1 if kn >= 2 else n + 2
    q = np.zeros((max_steps + 2) * n, dtype=np.float64)
    h = np.zeros((max_steps + 2) * (max_steps + 1), dtype=np.float64)
    ncq, nmax = C.c_int(0), C.c_int(0)
    rc = lib().ora_arnoldi(aa._h, b._h, kn, max_steps, q.ctypes.data_as(_pf64), h.ctypes.data_as(_pf64),
                           C.byref(ncq), C.byref(nmax))
    if rc:
        raise OracleError(rc, "arnoldi")
    Q = q[: ncq.value * n].reshape(ncq.value, n).T.copy()
    H = h[: (nmax.value + 1) * nmax.value].reshape(nmax.value, nmax.value + 1).T.copy()
    return Q, H


def synth_row(kind, n, k, seed, band, i):
    cols = np.zeros(128, dtype=np.int64)
    vals = np.zeros(128, dtype=np.float64)
    cnt = C.c_int(0)
    lib().ora_synth_row(kind, n, k, seed, band, i, cols.ctypes.data_as(_pi64), vals.ctypes.data_as(_pf64),
                        C.byref(cnt))
    return cols[: cnt.value].copy(), vals[: cnt.value].copy()


def time_matvec(aa, x, reps=3, threads=1):
    cs = C.c_double(0)
    return lib().ora_time_matvec(aa._h, x._h, reps, threads, C.byref(cs))


def time_bicgstab(aa, b, x0, steps=3):
    cs = C.c_double(0)
    return lib().ora_time_bicgstab(aa._h, b._h, x0._h, steps, C.byref(cs))
