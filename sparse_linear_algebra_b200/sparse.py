"""Host-side mirror of Numeric.LinearAlgebra.Sparse's operator surface over libsla_b200.so.

The reference is Haskell and no GHC exists in this image, so the host side above the C ABI is written in
Python with the reference's names, argument order and error behaviour (the Haskell shim a maintainer would
add is in hs/ and INTEGRATION.md).  Everything here is marshalling: all arithmetic runs in the CUDA library.

    reference (src/Numeric/LinearAlgebra/...)         here
    aa #> v          Class.hs:224-229                  aa.matVec(v)   or  aa @ v
    v <# aa                                            aa.vecMat(v)
    v <.> w          Class.hs:81-87                    v.dot(w)
    v ^+^ w, v ^-^ w Class.hs:57-69                    v + w, v - w
    a .* v, v ./ s   Class.hs:72-95                    a * v, v / s
    norm2, normalize2  Class.hs:126-153                v.norm2(), v.normalize2()
    transpose aa     Class.hs:195-207                  aa.transpose()
    bicgsInit / bicgstabStep   Sparse.hs:965-981       bicgsInit(aa, b, x0) / bicgstabStep(aa, r0hat, st)
    cgsInit / cgsStep          Sparse.hs:923-939       cgsInit / cgsStep
    cgneInit / cgneStep        Sparse.hs:862-878       cgneInit / cgneStep
    linSolve0 method aa b x0   Sparse.hs:1016-1072     linSolve0(method, aa, b, x0)
    arnoldi aa b kn            Sparse.hs:630-667       arnoldi(aa, b, kn)
    aa <\\> b                   Class.hs:244-249        backslash(aa, b)   (GMRES(30), x0 = 0.1, Sparse.hs:1082-1088)

Krylov steps advance the state record IN PLACE on the device and return it (the reference returns a new
immutable record); keep a copy of the fields (st.x.toDenseListSV()) if the old state is needed.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L

GMRES_, CGNE_, BCG_, CGS_, BICGSTAB_ = 0, 1, 2, 3, 4          # LinSolveMethod, Sparse.hs:1007-1012
GEN_UNIFORM, GEN_BANDED, GEN_LAPLACE2D, GEN_BLOCK16 = 0, 1, 2, 3              # include/sla_synth.h


class SlaError(Exception):
    """Base of the errors raised by the backend; `.status` is the sla_status code."""

    def __init__(self, status, message):
        super().__init__(f"{L.STATUS_NAMES[status] if 0 <= status < len(L.STATUS_NAMES) else status}: {message}")
        self.status = status
        self.message = message


class MatVecSizeMismatchException(SlaError):
    """OperandSizeMismatch / `error "matVec : mismatched dimensions"` (Control/Exception/Common.hs:44-51)."""


class OutOfBoundsIndexError(SlaError):
    """`error "insertSpMatrix : index out of bounds"` (SpMatrix.hs:205-208)."""


class IterE(SlaError):
    """IterationException IterE (Control/Exception/Common.hs:67-76), e.g. unsupported linSolve0 method."""


class NeedsPivoting(SlaError):
    """MatrixException NeedsPivoting: a nearZero diagonal met by a triangular solve (Control/Exception/Common.hs:57-61)."""


_ERR_CLASS = {L.SLA_ERR_SIZE_MISMATCH: MatVecSizeMismatchException, L.SLA_ERR_OOB_INDEX: OutOfBoundsIndexError,
              L.SLA_ERR_UNSUPPORTED_METHOD: IterE, L.SLA_ERR_NEEDS_PIVOTING: NeedsPivoting}


class Context:
    """One GPU (one sla_ctx).  Not thread-safe, like the single-threaded reference."""

    def __init__(self, device=None, rank=0, world=1, nccl_id=None):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        if world > 1:
            buf = C.create_string_buffer(bytes(nccl_id), 128)
            st = self.lib.sla_init_dist(device, rank, world, buf, C.byref(h))
        else:
            st = self.lib.sla_init(device, C.byref(h))
        if st != L.SLA_OK:
            raise SlaError(st, (self.lib.sla_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def check(self, st, ok=(L.SLA_OK,)):
        if st in ok:
            return st
        msg = (self.lib.sla_last_error(self.h) or b"").decode()
        raise _ERR_CLASS.get(st, SlaError)(st, msg)

    def sync(self):
        self.check(self.lib.sla_sync(self.h))

    def set_option(self, name, value):
        self.check(self.lib.sla_set_option(self.h, name.encode(), int(value)))

    @property
    def launches(self):
        return self.lib.sla_launch_count(self.h)

    @property
    def stream(self):
        return self.lib.sla_stream(self.h)

    def timer_start(self):
        self.check(self.lib.sla_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self.check(self.lib.sla_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def pinned(self, n, dtype=np.float64):
        """A page-locked numpy array of n elements (freed with the returned object's `.free()` or at exit)."""
        nbytes = int(n) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        self.check(self.lib.sla_host_alloc(self.h, nbytes, C.byref(ptr)))
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        return arr

    def close(self):
        if getattr(self, "h", None):
            self.lib.sla_finalize(self.h)
            self.h = None


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


class SpVector:
    """SpVector Double on the device: a dense double[n]; absent keys of the reference's IntMap are 0.0."""

    def __init__(self, ctx, handle, owns=True):
        self.ctx, self.h, self._owns = ctx, handle, owns

    def __del__(self):
        if getattr(self, "_owns", False) and getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.sla_vec_free(self.h)
            self.h = None

    # -- construction (SpVector.hs:157-195, 232-233, 275-278)
    @staticmethod
    def mkSpVR(d, ll, ctx=None):
        ctx = ctx or default_context()
        x = np.zeros(d, dtype=np.float64)
        ll = np.asarray(ll, dtype=np.float64)[:d]
        x[: ll.size] = ll
        a, p = _f64(x)
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_vec_from_host(ctx.h, d, p, C.byref(h)))
        return SpVector(ctx, h)

    fromListDenseSV = mkSpVR

    @staticmethod
    def fromListSV(d, iix, ctx=None):
        # foldr insert: the FIRST occurrence of an index wins, out-of-bounds entries are dropped
        x = np.zeros(d, dtype=np.float64)
        for i, v in reversed(list(iix)):
            if 0 <= i < d:
                x[i] = v
        return SpVector.mkSpVR(d, x, ctx)

    @staticmethod
    def zeroSV(d, ctx=None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_vec_create(ctx.h, d, C.byref(h)))
        return SpVector(ctx, h)

    @staticmethod
    def constv(d, x, ctx=None):
        v = SpVector.zeroSV(d, ctx)
        v.ctx.check(v.ctx.lib.sla_vec_fill(v.ctx.h, v.h, float(x)))
        return v

    @staticmethod
    def onesSV(d, ctx=None):
        return SpVector.constv(d, 1.0, ctx)

    @staticmethod
    def generate(n, seed, ctx=None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_vec_generate(ctx.h, n, seed, C.byref(h)))
        return SpVector(ctx, h)

    def _like(self):
        return SpVector.zeroSV(self.dim, self.ctx)

    def copy(self):
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_copy(self.ctx.h, self.h, z.h))
        return z

    # -- inspection
    @property
    def dim(self):
        return self.ctx.lib.sla_vec_dim(self.h)

    def toDenseListSV(self):
        out = np.zeros(self.dim, dtype=np.float64)
        self.ctx.check(self.ctx.lib.sla_vec_to_host(self.ctx.h, self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    # -- algebra
    def __add__(self, w):
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_add(self.ctx.h, self.h, w.h, z.h))
        return z

    def __sub__(self, w):
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_sub(self.ctx.h, self.h, w.h, z.h))
        return z

    def __neg__(self):
        return (-1.0) * self           # negateV: IEEE negation == multiplication by -1

    def __rmul__(self, a):
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_scale(self.ctx.h, float(a), self.h, z.h))
        return z

    def __truediv__(self, s):          # v ./ s = recip s .* v
        return (1.0 / float(s)) * self

    def axpy(self, a, x):
        """self ^+^ (a .* x), rounded as written (one fused kernel)."""
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_axpy(self.ctx.h, float(a), x.h, self.h, z.h))
        return z

    def dot(self, w):
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.sla_dot(self.ctx.h, self.h, w.h, C.byref(out)))
        return out.value

    def norm2Sq(self):
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.sla_norm2sq(self.ctx.h, self.h, C.byref(out)))
        return out.value

    def norm2(self):
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.sla_norm2(self.ctx.h, self.h, C.byref(out)))
        return out.value

    def normalize2(self):
        z = self._like()
        self.ctx.check(self.ctx.lib.sla_vec_normalize2(self.ctx.h, self.h, z.h))
        return z


class DenseBlock:
    """Dense column-major block on the device (the Krylov basis Q of `arnoldi`)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.sla_dense_free(self.h)
            self.h = None

    @property
    def dim(self):
        r, c = C.c_int64(0), C.c_int64(0)
        self.ctx.check(self.ctx.lib.sla_dense_dims(self.h, C.byref(r), C.byref(c)))
        return r.value, c.value

    def column(self, j):
        """Column j as a new SpVector (extractCol of the reference's Q)."""
        v = C.c_void_p()
        self.ctx.check(self.ctx.lib.sla_dense_column(self.ctx.h, self.h, j, C.byref(v)))
        return SpVector(self.ctx, v)

    def toHost(self):
        r, c = self.dim
        out = np.zeros((c, r), dtype=np.float64)
        self.ctx.check(self.ctx.lib.sla_dense_to_host(self.ctx.h, self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.T.copy()


F64, BF16 = 0, 1


class DenseMatrix(DenseBlock):
    """Row-major dense block: the dense right operand / result of (##) (fp64 or bf16)."""

    def __init__(self, ctx, handle, dtype):
        super().__init__(ctx, handle)
        self.dtype = dtype

    @staticmethod
    def fromHost(a, dtype=F64, ctx=None):
        ctx = ctx or default_context()
        a = np.ascontiguousarray(a, dtype=np.float64)
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_dense_from_host(ctx.h, a.shape[0], a.shape[1], a.ctypes.data_as(C.POINTER(C.c_double)), dtype, C.byref(h)))
        return DenseMatrix(ctx, h, dtype)

    @staticmethod
    def generate(rows, cols, seed, dtype=F64, ctx=None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_dense_generate(ctx.h, rows, cols, seed, dtype, C.byref(h)))
        return DenseMatrix(ctx, h, dtype)

    @staticmethod
    def zeros(rows, cols, dtype=F64, ctx=None):
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_dense_create(ctx.h, rows, cols, dtype, C.byref(h)))
        return DenseMatrix(ctx, h, dtype)

    def toHost(self):
        r, c = self.dim
        out = np.zeros((r, c), dtype=np.float64)
        self.ctx.check(self.ctx.lib.sla_dense_to_host_f64(self.ctx.h, self.h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out


class SpMatrix:
    """SpMatrix Double on the device: CSR (row_ptr, col_idx ascending, val)."""

    def __init__(self, ctx, handle, owns=True):
        self.ctx, self.h, self._owns = ctx, handle, owns

    def __del__(self):
        if getattr(self, "_owns", False) and getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.sla_csr_free(self.h)
            self.h = None

    # -- construction (SpMatrix.hs:128-135, 184-191, 205-241)
    @staticmethod
    def fromCOO(dims, i, j, v, ctx=None):
        ia, ip = _i64(i)
        ja, jp = _i64(j)
        va, vp = _f64(v)
        if not (ia.size == ja.size == va.size):          # the C side would read past the shorter buffer
            raise ValueError(f"fromCOO: i, j and v must have the same length ({ia.size}, {ja.size}, {va.size})")
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_csr_from_coo(ctx.h, dims[0], dims[1], ia.size, ip, jp, vp, C.byref(h)))
        return SpMatrix(ctx, h)

    @staticmethod
    def fromListSM(dims, iix, ctx=None):
        iix = list(iix)
        return SpMatrix.fromCOO(dims, [t[0] for t in iix], [t[1] for t in iix], [t[2] for t in iix], ctx)

    @staticmethod
    def fromListDenseSM(m, ll, ctx=None):
        ll = list(ll)
        n = len(ll) // m if m else 0            # column-major, truncated to n*m elements
        q = np.arange(n * m)
        return SpMatrix.fromCOO((m, n), q % m if m else q, q // m if m else q, ll[: n * m], ctx)

    @staticmethod
    def fromCSR(m, n, row_ptr, col_idx, val, ctx=None):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        va, vp = _f64(val)
        if rp.size != m + 1 or ci.size != va.size:
            raise ValueError(f"fromCSR: row_ptr needs m + 1 = {m + 1} entries (got {rp.size}) and col_idx / val equal lengths ({ci.size}, {va.size})")
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_csr_from_csr(ctx.h, m, n, ci.size, rp.ctypes.data_as(C.POINTER(C.c_int32)),
                                           ci.ctypes.data_as(C.POINTER(C.c_int32)), vp, C.byref(h)))
        return SpMatrix(ctx, h)

    @staticmethod
    def fromMatrixMarket(path, ctx=None):
        """Coordinate real Matrix Market file -> device matrix (fromListSM semantics for duplicates)."""
        from .mmio import read_matrix_market

        m, n, i, j, v = read_matrix_market(path)
        return SpMatrix.fromCOO((m, n), i, j, v, ctx)

    @staticmethod
    def eye(n, ctx=None):
        return SpMatrix.mkDiagonal(n, np.ones(n), ctx)

    @staticmethod
    def mkDiagonal(n, xx, ctx=None):
        if len(xx) < n:
            raise ValueError(f"mkDiagonal: {n} diagonal entries expected, got {len(xx)}")
        return SpMatrix.fromCOO((n, n), np.arange(n), np.arange(n), np.asarray(xx, dtype=np.float64)[:n], ctx)

    @staticmethod
    def generate(kind, n, nnz_per_row, seed, band=0, ctx=None):
        """Synthetic workloads of SURVEY.md §8(d), generated on the device (include/sla_synth.h)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        ctx.check(ctx.lib.sla_csr_generate(ctx.h, kind, n, nnz_per_row, seed, band, C.byref(h)))
        return SpMatrix(ctx, h)

    # -- inspection
    def _dims(self):
        m, n, z = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.ctx.check(self.ctx.lib.sla_csr_dims(self.h, C.byref(m), C.byref(n), C.byref(z)))
        return m.value, n.value, z.value

    nrows = property(lambda s: s._dims()[0])
    ncols = property(lambda s: s._dims()[1])
    nnz = property(lambda s: s._dims()[2])
    dim = property(lambda s: s._dims()[:2])

    @property
    def spmv_bytes(self):
        return self.ctx.lib.sla_csr_spmv_bytes(self.h)

    def toCSR(self):
        m, n, z = self._dims()
        rp = np.zeros(m + 1, dtype=np.int32)
        ci = np.zeros(z, dtype=np.int32)
        va = np.zeros(z, dtype=np.float64)
        self.ctx.check(self.ctx.lib.sla_csr_to_host(self.ctx.h, self.h, rp.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    ci.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    va.ctypes.data_as(C.POINTER(C.c_double))))
        return rp, ci, va

    def toDense(self):
        m, n, _ = self._dims()
        rp, ci, va = self.toCSR()
        d = np.zeros((m, n))
        for r in range(m):
            d[r, ci[rp[r]: rp[r + 1]]] = va[rp[r]: rp[r + 1]]
        return d

    def isDiagonalSM(self):
        out = C.c_int(0)
        self.ctx.check(self.ctx.lib.sla_csr_is_diagonal(self.ctx.h, self.h, C.byref(out)))
        return bool(out.value)

    # -- algebra
    def transpose(self):
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sla_csr_transpose(self.ctx.h, self.h, C.byref(h)))
        return SpMatrix(self.ctx, h)

    def matVec(self, x, out=None):        # aa #> x
        y = out if out is not None else SpVector.zeroSV(self.nrows, self.ctx)
        self.ctx.check(self.ctx.lib.sla_spmv(self.ctx.h, self.h, x.h, y.h))
        return y

    def vecMat(self, x, out=None):        # x <# aa
        y = out if out is not None else SpVector.zeroSV(self.ncols, self.ctx)
        self.ctx.check(self.ctx.lib.sla_spmvT(self.ctx.h, self.h, x.h, y.h))
        return y

    def matVecHost(self, x_host):
        """(#>) on host buffers: H2D copy, kernel, D2H copy inside one C-ABI call."""
        m, n, _ = self._dims()
        xa, xp = _f64(x_host)
        if xa.size != n:
            raise MatVecSizeMismatchException(L.SLA_ERR_SIZE_MISMATCH, f"matVec : mismatched dimensions ({n},{xa.size})")
        y = np.empty(m, dtype=np.float64)
        self.ctx.check(self.ctx.lib.sla_spmv_host(self.ctx.h, self.h, xp, y.ctypes.data_as(C.POINTER(C.c_double))))
        return y

    # -- diagonal partitions (SpMatrix.hs:306-315) and the schedule of the triangular solves
    def _parts(self):
        e, d, f = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.ctx.check(self.ctx.lib.sla_csr_diag_partitions(self.ctx.h, self.h, C.byref(e), C.byref(d), C.byref(f)))
        return SpMatrix(self.ctx, e), SpMatrix(self.ctx, d), SpMatrix(self.ctx, f)

    def extractSubDiag(self):
        return self._parts()[0]

    def extractDiag(self):
        return self._parts()[1]

    def extractSuperDiag(self):
        return self._parts()[2]

    def triAnalysis(self, upper=False):
        """(dependency levels, stored entries of the triangle incl. the diagonal) of the cached level schedule."""
        lv, nz = C.c_int(0), C.c_int64(0)
        self.ctx.check(self.ctx.lib.sla_tri_analysis(self.ctx.h, self.h, 1 if upper else 0, C.byref(lv), C.byref(nz)))
        return lv.value, nz.value

    def matMat(self, b, out=None):        # aa ## b, b a dense row-major block
        cc = out if out is not None else DenseMatrix.zeros(self.nrows, b.dim[1], b.dtype, self.ctx)
        self.ctx.check(self.ctx.lib.sla_spmm_dense(self.ctx.h, self.h, b.h, cc.h))
        return cc

    def matMatT(self, bt, out=None):      # aa ##^ b  = aa ## transpose b   (Class.hs:199-200); bt holds b itself, k x n row-major
        cc = out if out is not None else DenseMatrix.zeros(self.nrows, bt.dim[0], bt.dtype, self.ctx)
        self.ctx.check(self.ctx.lib.sla_spmm_dense_abt(self.ctx.h, self.h, bt.h, cc.h))
        return cc

    def tMatMat(self, b, out=None):       # aa #^# b  = transpose aa ## b   (Class.hs:201-203)
        cc = out if out is not None else DenseMatrix.zeros(self.ncols, b.dim[1], b.dtype, self.ctx)
        self.ctx.check(self.ctx.lib.sla_spmm_dense_atb(self.ctx.h, self.h, b.h, cc.h))
        return cc

    def normFrobenius(self):              # sqrt (trace (m ##^ m))   SpMatrix.hs:751-752
        out = C.c_double(0)
        self.ctx.check(self.ctx.lib.sla_csr_norm_frobenius(self.ctx.h, self.h, C.byref(out)))
        return out.value

    def __matmul__(self, x):
        return self.matMat(x) if isinstance(x, DenseMatrix) else self.matVec(x)


class KrylovState:
    """BICGSTAB / CGS / CGNE record living on the device; fields are borrowed views."""

    def __init__(self, ctx, handle, kind):
        self.ctx, self.h, self.kind = ctx, handle, kind

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.sla_krylov_free(self.h)
            self.h = None

    def clone(self):
        """Deep copy of the record: what a pure `step` (the reference's signature) advances instead of its argument."""
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.sla_krylov_clone(self.ctx.h, self.h, C.byref(h)))
        return KrylovState(self.ctx, h, self.kind)

    def _field(self, f):
        v = C.c_void_p()
        self.ctx.check(self.ctx.lib.sla_krylov_view(self.ctx.h, self.h, f, C.byref(v)))
        view = SpVector(self.ctx, v, owns=False)
        view._keepalive = self
        return view

    x = property(lambda s: s._field(0))     # _x / _xBicgstab / _xCgne
    r = property(lambda s: s._field(1))
    p = property(lambda s: s._field(2))
    u = property(lambda s: s._field(3))


def _init(fn_name, kind, aa, b, x0):
    ctx = aa.ctx
    h = C.c_void_p()
    ctx.check(getattr(ctx.lib, fn_name)(ctx.h, aa.h, b.h, x0.h, C.byref(h)))
    return KrylovState(ctx, h, kind)


def bicgsInit(aa, b, x0):
    return _init("sla_bicgstab_init", BICGSTAB_, aa, b, x0)


def bicgstabStep(aa, r0hat, st, pure=False):
    """bicgstabStep aa r0hat st (Sparse.hs:970-981).  pure=True is the reference's signature (st stays valid, a new record is
    returned: `iterate (bicgstabStep aa r0hat) st0 !! 20`, README.md:208); the default advances st in place."""
    if pure:
        st = st.clone()
    aa.ctx.check(aa.ctx.lib.sla_bicgstab_step(aa.ctx.h, aa.h, r0hat.h, st.h))
    return st


def cgsInit(aa, b, x0):
    return _init("sla_cgs_init", CGS_, aa, b, x0)


def cgsStep(aa, rhat, st, pure=False):
    if pure:
        st = st.clone()
    aa.ctx.check(aa.ctx.lib.sla_cgs_step(aa.ctx.h, aa.h, rhat.h, st.h))
    return st


def cgneInit(aa, b, x0):
    return _init("sla_cgne_init", CGNE_, aa, b, x0)


def cgneStep(aa, st, pure=False):
    if pure:
        st = st.clone()
    aa.ctx.check(aa.ctx.lib.sla_cgne_step(aa.ctx.h, aa.h, st.h))
    return st


def _opts(nits, tol_abs, tol_rel, true_residual, check_every):
    o = L.SolveOpts()
    L.load().sla_solve_opts_default(C.byref(o))
    if nits:
        o.max_iters = nits
    if tol_abs:
        o.tol_abs = tol_abs
    if tol_rel:
        o.tol_rel = tol_rel
    o.true_residual = 1 if true_residual else 0
    o.check_every = check_every
    return o


def linSolve0(method, aa, b, x0, nits=0, tol_abs=0.0, tol_rel=0.0, true_residual=True, check_every=1, info=False):
    """linSolve0 method aa b x0 (Sparse.hs:1016-1072): nits = 200, tol = max 1e-6 (1e-4 * ||r0||), true residual."""
    ctx = aa.ctx
    x = SpVector.zeroSV(x0.dim, ctx)
    o = _opts(nits, tol_abs, tol_rel, true_residual, check_every)
    iters, res = C.c_int(0), C.c_double(0)
    ctx.check(ctx.lib.sla_linsolve0(ctx.h, method, aa.h, b.h, x0.h, C.byref(o), x.h, C.byref(iters), C.byref(res)))
    return (x, iters.value, res.value) if info else x


def linSolve0Host(method, aa, b_host, x0_host, nits=0, tol_abs=0.0, tol_rel=0.0, info=False):
    ctx = aa.ctx
    ba, bp = _f64(b_host)
    xa, xp = _f64(x0_host)
    out = np.empty(aa.ncols, dtype=np.float64)
    o = _opts(nits, tol_abs, tol_rel, True, 1)
    iters, res = C.c_int(0), C.c_double(0)
    ctx.check(ctx.lib.sla_linsolve0_host(ctx.h, method, aa.h, bp, xp, C.byref(o), out.ctypes.data_as(C.POINTER(C.c_double)),
                                         C.byref(iters), C.byref(res)))
    return (out, iters.value, res.value) if info else out


def arnoldi(aa, b, kn):
    """arnoldi aa b kn (Sparse.hs:630-667) -> (Q DenseBlock n x (nmax+1), H numpy (nmax+1) x nmax, breakdown flag)."""
    ctx = aa.ctx
    h = np.zeros((kn + 1) * kn, dtype=np.float64)
    q = C.c_void_p()
    nmax = C.c_int(0)
    st = ctx.check(ctx.lib.sla_arnoldi(ctx.h, aa.h, b.h, kn, C.byref(q), h.ctypes.data_as(C.POINTER(C.c_double)),
                                       C.byref(nmax)), ok=(L.SLA_OK, L.SLA_ERR_BREAKDOWN))
    k = nmax.value
    H = h[: (k + 1) * k].reshape(k, k + 1).T.copy()
    return DenseBlock(ctx, q), H, st == L.SLA_ERR_BREAKDOWN


def gmres(aa, b, x0, restart=30, nits=0, tol_abs=0.0, tol_rel=0.0, info=False, fixed_work=False):
    """Restarted GMRES(restart).  fixed_work=True runs exactly `nits` Arnoldi steps (no stopping test; BASELINE config 4's
    "GMRES(30), 10 restarts" is restart=30, nits=300)."""
    ctx = aa.ctx
    x = SpVector.zeroSV(x0.dim, ctx)
    o = _opts(nits, tol_abs, tol_rel, True, -1 if fixed_work else 1)
    iters, res = C.c_int(0), C.c_double(0)
    ctx.check(ctx.lib.sla_gmres(ctx.h, aa.h, b.h, x0.h, restart, C.byref(o), x.h, C.byref(iters), C.byref(res)))
    return (x, iters.value, res.value) if info else x


def diagPartitions(aa):
    """diagPartitions aa = (e, d, f): strictly sub-diagonal, diagonal, strictly super-diagonal parts (Sparse.hs:673-679)."""
    return aa._parts()


def jacobiPre(aa):
    """jacobiPre x = recip <$> extractDiag x (Sparse.hs:686-687)."""
    h = C.c_void_p()
    aa.ctx.check(aa.ctx.lib.sla_jacobi_pre(aa.ctx.h, aa.h, C.byref(h)))
    return SpMatrix(aa.ctx, h)


def mSsorPre(aa, omega):
    """mSsorPre aa omega = (l, r), l = (eye n ^-^ scale omega e) ## reciprocal d, r = d ^-^ scale omega f (Sparse.hs:713-721)."""
    l, r = C.c_void_p(), C.c_void_p()
    aa.ctx.check(aa.ctx.lib.sla_mssor_pre(aa.ctx.h, aa.h, float(omega), C.byref(l), C.byref(r)))
    return SpMatrix(aa.ctx, l), SpMatrix(aa.ctx, r)


def ilu0Pre(aa):
    """ilu0Pre aa = (l, u) with holes: the reference's complete `lu` masked by aa's stored positions (Sparse.hs:696-706)."""
    l, u = C.c_void_p(), C.c_void_p()
    aa.ctx.check(aa.ctx.lib.sla_ilu0_pre(aa.ctx.h, aa.h, C.byref(l), C.byref(u)))
    return SpMatrix(aa.ctx, l), SpMatrix(aa.ctx, u)


def triLowerSolve(ll, b, out=None):
    """triLowerSolve ll b: forward substitution (Sparse.hs:750-777); raises NeedsPivoting on a nearZero diagonal."""
    w = out if out is not None else SpVector.zeroSV(b.dim, ll.ctx)
    ll.ctx.check(ll.ctx.lib.sla_tri_lower_solve(ll.ctx.h, ll.h, b.h, w.h))
    return w


def triUpperSolve(uu, w, out=None):
    """triUpperSolve uu w: backward substitution (Sparse.hs:784-811)."""
    x = out if out is not None else SpVector.zeroSV(w.dim, uu.ctx)
    uu.ctx.check(uu.ctx.lib.sla_tri_upper_solve(uu.ctx.h, uu.h, w.h, x.h))
    return x


def backslash(aa, b):
    """aa <\\> b: the commented-out LinearSystem instance uses GMRES with x0 = 0.1 (Sparse.hs:1082-1088)."""
    return gmres(aa, b, SpVector.constv(b.dim, 0.1, aa.ctx), restart=30)
