"""One host process, several GPUs: the Python face of csrc/multi.cu (sla_init_multi).

    m = MultiContext(2)                      # GPUs 0 and 1 of this box, driven from this one process
    A = m.generate(GEN_LAPLACE2D, n, 5, 0, g)   # global, row-partitioned matrix
    x = m.vector(np.ones(n))                 # global vector (scattered)
    y = A @ x                                # (#>) : x exchange + kernels on both GPUs
    y.toDenseListSV()                        # gathered

Objects are GLOBAL: the caller never sees ranks.  The torchrun path (dist.py, one process per GPU) stays
for jobs that already run that way."""
import ctypes as C

import numpy as np

from . import _lib as L
from .sparse import _ERR_CLASS, SlaError, _opts

_pd = C.POINTER(C.c_double)


class MultiContext:
    def __init__(self, n_gpus, device_ids=None):
        self.lib = L.load()
        h = C.c_void_p()
        ids = (C.c_int * n_gpus)(*device_ids) if device_ids is not None else None
        st = self.lib.sla_init_multi(n_gpus, ids, C.byref(h))
        if st != L.SLA_OK:
            raise SlaError(st, (self.lib.sla_multi_last_error(None) or b"").decode())
        self.h = h
        self.world = n_gpus

    def check(self, st, ok=(L.SLA_OK,)):
        if st in ok:
            return st
        raise _ERR_CLASS.get(st, SlaError)(st, (self.lib.sla_multi_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.sla_finalize_multi(self.h)
            self.h = None

    @property
    def p2p(self):
        return all(self.lib.sla_p2p_enabled(self.lib.sla_multi_ctx(self.h, r)) for r in range(self.world)) if self.world > 1 else False

    @property
    def launches(self):
        return sum(self.lib.sla_launch_count(self.lib.sla_multi_ctx(self.h, r)) for r in range(self.world))

    # -- construction
    def generate(self, kind, n, nnz_per_row, seed, band=0):
        h = C.c_void_p()
        self.check(self.lib.sla_multi_csr_generate(self.h, kind, n, nnz_per_row, seed, band, C.byref(h)))
        return MultiMatrix(self, h)

    def fromCSR(self, m, n, row_ptr, col_idx, val):
        rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
        ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        va = np.ascontiguousarray(val, dtype=np.float64)
        if rp.size != m + 1 or ci.size != va.size:
            raise ValueError("fromCSR: row_ptr needs m + 1 entries and col_idx / val equal lengths")
        h = C.c_void_p()
        self.check(self.lib.sla_multi_csr_from_csr(self.h, m, n, ci.size, rp.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   ci.ctypes.data_as(C.POINTER(C.c_int32)), va.ctypes.data_as(_pd), C.byref(h)))
        return MultiMatrix(self, h)

    def vector(self, x):
        a = np.ascontiguousarray(x, dtype=np.float64)
        h = C.c_void_p()
        self.check(self.lib.sla_multi_vec_from_host(self.h, a.size, a.ctypes.data_as(_pd), C.byref(h)))
        return MultiVector(self, h)

    def zeros(self, n):
        h = C.c_void_p()
        self.check(self.lib.sla_multi_vec_create(self.h, n, C.byref(h)))
        return MultiVector(self, h)

    def generate_vector(self, n, seed):
        h = C.c_void_p()
        self.check(self.lib.sla_multi_vec_generate(self.h, n, seed, C.byref(h)))
        return MultiVector(self, h)


class MultiVector:
    def __init__(self, m, h):
        self.m, self.h = m, h

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.m, "h", None):
            self.m.lib.sla_multi_vec_free(self.h)
            self.h = None

    @property
    def dim(self):
        return self.m.lib.sla_multi_vec_dim(self.h)

    def toDenseListSV(self):
        out = np.zeros(self.dim)
        self.m.check(self.m.lib.sla_multi_vec_to_host(self.m.h, self.h, out.ctypes.data_as(_pd)))
        return out

    def copy(self):
        z = self.m.zeros(self.dim)
        self.m.check(self.m.lib.sla_multi_vec_copy(self.m.h, self.h, z.h))
        return z

    def dot(self, w):
        out = C.c_double(0)
        self.m.check(self.m.lib.sla_multi_dot(self.m.h, self.h, w.h, C.byref(out)))
        return out.value

    def norm2(self):
        out = C.c_double(0)
        self.m.check(self.m.lib.sla_multi_norm2(self.m.h, self.h, C.byref(out)))
        return out.value

    def axpy(self, a, x):                     # self ^+^ (a .* x)
        z = self.m.zeros(self.dim)
        self.m.check(self.m.lib.sla_multi_vec_axpy(self.m.h, float(a), x.h, self.h, z.h))
        return z

    def __rmul__(self, a):
        z = self.m.zeros(self.dim)
        self.m.check(self.m.lib.sla_multi_vec_scale(self.m.h, float(a), self.h, z.h))
        return z


class MultiMatrix:
    def __init__(self, m, h):
        self.m, self.h = m, h

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.m, "h", None):
            self.m.lib.sla_multi_csr_free(self.h)
            self.h = None

    @property
    def dim(self):
        r, c, z = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.m.check(self.m.lib.sla_multi_csr_dims(self.h, C.byref(r), C.byref(c), C.byref(z)))
        return r.value, c.value

    @property
    def nnz(self):
        z = C.c_int64(0)
        self.m.check(self.m.lib.sla_multi_csr_dims(self.h, None, None, C.byref(z)))
        return z.value

    def matVec(self, x, out=None):
        y = out if out is not None else self.m.zeros(self.dim[0])
        self.m.check(self.m.lib.sla_multi_spmv(self.m.h, self.h, x.h, y.h))
        return y

    __matmul__ = matVec


class MultiKrylov:
    def __init__(self, m, h):
        self.m, self.h = m, h

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.m, "h", None):
            self.m.lib.sla_multi_krylov_free(self.h)
            self.h = None

    def field(self, f, n):
        out = np.zeros(n)
        self.m.check(self.m.lib.sla_multi_krylov_get(self.m.h, self.h, f, out.ctypes.data_as(_pd)))
        return out

    def clone(self):
        h = C.c_void_p()
        self.m.check(self.m.lib.sla_multi_krylov_clone(self.m.h, self.h, C.byref(h)))
        return MultiKrylov(self.m, h)


def bicgsInit(aa, b, x0):
    h = C.c_void_p()
    aa.m.check(aa.m.lib.sla_multi_bicgstab_init(aa.m.h, aa.h, b.h, x0.h, C.byref(h)))
    return MultiKrylov(aa.m, h)


def bicgstabStep(aa, r0hat, st, pure=False):
    if pure:
        st = st.clone()
    aa.m.check(aa.m.lib.sla_multi_bicgstab_step(aa.m.h, aa.h, r0hat.h, st.h))
    return st


def linSolve0(method, aa, b, x0, nits=0, tol_abs=0.0, tol_rel=0.0, info=False):
    x = aa.m.zeros(x0.dim)
    o = _opts(nits, tol_abs, tol_rel, True, 1)
    it, res = C.c_int(0), C.c_double(0)
    aa.m.check(aa.m.lib.sla_multi_linsolve0(aa.m.h, method, aa.h, b.h, x0.h, C.byref(o), x.h, C.byref(it), C.byref(res)))
    return (x, it.value, res.value) if info else x


def gmres(aa, b, x0, restart=30, nits=0, tol_abs=0.0, tol_rel=0.0, info=False):
    x = aa.m.zeros(x0.dim)
    o = _opts(nits, tol_abs, tol_rel, True, 1)
    it, res = C.c_int(0), C.c_double(0)
    aa.m.check(aa.m.lib.sla_multi_gmres(aa.m.h, aa.h, b.h, x0.h, restart, C.byref(o), x.h, C.byref(it), C.byref(res)))
    return (x, it.value, res.value) if info else x


def arnoldi(aa, b, kn):
    """(Q as a numpy n x (nmax+1) array gathered from the GPUs, H numpy (nmax+1) x nmax, breakdown flag)."""
    m = aa.m
    h = np.zeros((kn + 1) * kn)
    q, nmax = C.c_void_p(), C.c_int(0)
    st = m.check(m.lib.sla_multi_arnoldi(m.h, aa.h, b.h, kn, C.byref(q), h.ctypes.data_as(_pd), C.byref(nmax)), ok=(L.SLA_OK, L.SLA_ERR_BREAKDOWN))
    k = nmax.value
    n = aa.dim[0]
    Q = np.zeros((k + 1, n))
    m.check(m.lib.sla_multi_dense_to_host(m.h, q, Q.ctypes.data_as(_pd)))
    m.lib.sla_multi_dense_free(q)
    return Q.T.copy(), h[: (k + 1) * k].reshape(k, k + 1).T.copy(), st == L.SLA_ERR_BREAKDOWN
