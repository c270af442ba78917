"""ctypes binding of libsla_b200.so (the C ABI declared in include/sla_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import, and if no CUDA
device is present `sla_init` fails with SLA_ERR_CUDA.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLA_LIB_PATH") or os.path.join(_HERE, "libsla_b200.so")   # SLA_LIB_PATH: tuning experiments only

(SLA_OK, SLA_ERR_SIZE_MISMATCH, SLA_ERR_OOB_INDEX, SLA_ERR_UNSUPPORTED_METHOD, SLA_ERR_NOT_CONVERGED,
 SLA_ERR_BREAKDOWN, SLA_ERR_CUDA, SLA_ERR_COMM, SLA_ERR_ALLOC, SLA_ERR_INVALID, SLA_ERR_NEEDS_PIVOTING) = range(11)

STATUS_NAMES = ["SLA_OK", "SLA_ERR_SIZE_MISMATCH", "SLA_ERR_OOB_INDEX", "SLA_ERR_UNSUPPORTED_METHOD",
                "SLA_ERR_NOT_CONVERGED", "SLA_ERR_BREAKDOWN", "SLA_ERR_CUDA", "SLA_ERR_COMM", "SLA_ERR_ALLOC",
                "SLA_ERR_INVALID", "SLA_ERR_NEEDS_PIVOTING"]


class SolveOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tol_abs", C.c_double), ("tol_rel", C.c_double),
                ("true_residual", C.c_int), ("check_every", C.c_int)]


_p = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_i64 = C.c_int64
_f64 = C.c_double
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)
_pf64 = C.POINTER(C.c_double)
_pint = C.POINTER(C.c_int)
_popts = C.POINTER(SolveOpts)

# name -> (restype, argtypes); must list EVERY symbol include/sla_b200.h declares (tests check this)
SIGNATURES = {
    "sla_init": (C.c_int, [C.c_int, _pp]),
    "sla_nccl_unique_id": (C.c_int, [_p]),
    "sla_init_dist": (C.c_int, [C.c_int, C.c_int, C.c_int, _p, _pp]),
    "sla_finalize": (None, [_p]),
    "sla_last_error": (C.c_char_p, [_p]),
    "sla_version": (C.c_char_p, []),
    "sla_sync": (C.c_int, [_p]),
    "sla_stream": (_p, [_p]),
    "sla_rank": (C.c_int, [_p]),
    "sla_world": (C.c_int, [_p]),
    "sla_launch_count": (_i64, [_p]),
    "sla_set_option": (C.c_int, [_p, C.c_char_p, _i64]),
    "sla_host_alloc": (C.c_int, [_p, _i64, _pp]),
    "sla_host_free": (None, [_p]),
    "sla_timer_start": (C.c_int, [_p]),
    "sla_timer_stop": (C.c_int, [_p, C.POINTER(C.c_float)]),
    "sla_csr_from_coo": (C.c_int, [_p, _i64, _i64, _i64, _pi64, _pi64, _pf64, _pp]),
    "sla_csr_from_csr": (C.c_int, [_p, _i64, _i64, _i64, _pi32, _pi32, _pf64, _pp]),
    "sla_csr_generate": (C.c_int, [_p, C.c_int, _i64, C.c_int, C.c_uint64, _i64, _pp]),
    "sla_csr_generate_rows": (C.c_int, [_p, C.c_int, _i64, C.c_int, C.c_uint64, _i64, _i64, _i64, _pp]),
    "sla_csr_col_range": (C.c_int, [_p, _p, _pi64, _pi64]),
    "sla_csr_set_dist": (C.c_int, [_p, _p, _i64, C.c_int, _pint, _pint, _pi64, _pi64, C.c_int]),
    "sla_csr_set_halo": (C.c_int, [_p, _p, C.c_int, _pi64, _i64]),
    "sla_csr_transpose_dist": (C.c_int, [_p, _p, _pi64, _pp]),
    "sla_csr_attach_transpose": (C.c_int, [_p, _p, _p]),
    "sla_p2p_export": (C.c_int, [_p, _p]),
    "sla_p2p_attach": (C.c_int, [_p, _p]),
    "sla_p2p_enable": (C.c_int, [_p, C.c_int]),
    "sla_p2p_enabled": (C.c_int, [_p]),
    "sla_csr_p2p_export": (C.c_int, [_p, _p, _p]),
    "sla_csr_p2p_attach": (C.c_int, [_p, _p, _p]),
    "sla_csr_p2p_enable": (C.c_int, [_p, _p, C.c_int]),
    "sla_csr_p2p_mode": (C.c_int, [_p]),
    "sla_p2p_phase_schedule": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_int)]),
    "sla_csr_debug_rot_panels": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_char_p]),
    "sla_csr_npanels": (C.c_int, [_p]),
    "sla_debug_bsell_host": (C.c_int, [C.c_int, _i64, _p, _p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    "sla_debug_push_create": (C.c_int, [_p, _p, _p, C.POINTER(_p)]),
    "sla_debug_push_start": (C.c_int, [_p, _p, C.c_int, C.c_int]),
    "sla_debug_push_join": (C.c_int, [_p, _p]),
    "sla_debug_push_free": (None, [_p]),
    "sla_vec_generate_slice": (C.c_int, [_p, _i64, _i64, C.c_uint64, _pp]),
    "sla_csr_dims": (C.c_int, [_p, _pi64, _pi64, _pi64]),
    "sla_csr_to_host": (C.c_int, [_p, _p, _pi32, _pi32, _pf64]),
    "sla_csr_transpose": (C.c_int, [_p, _p, _pp]),
    "sla_csr_is_diagonal": (C.c_int, [_p, _p, _pint]),
    "sla_csr_spmv_bytes": (_i64, [_p]),
    "sla_csr_free": (None, [_p]),
    "sla_vec_create": (C.c_int, [_p, _i64, _pp]),
    "sla_vec_from_host": (C.c_int, [_p, _i64, _pf64, _pp]),
    "sla_vec_generate": (C.c_int, [_p, _i64, C.c_uint64, _pp]),
    "sla_vec_upload": (C.c_int, [_p, _p, _pf64]),
    "sla_vec_to_host": (C.c_int, [_p, _p, _pf64]),
    "sla_vec_copy": (C.c_int, [_p, _p, _p]),
    "sla_vec_fill": (C.c_int, [_p, _p, _f64]),
    "sla_vec_dim": (_i64, [_p]),
    "sla_vec_free": (None, [_p]),
    "sla_spmv": (C.c_int, [_p, _p, _p, _p]),
    "sla_spmvT": (C.c_int, [_p, _p, _p, _p]),
    "sla_dot": (C.c_int, [_p, _p, _p, _pf64]),
    "sla_norm2sq": (C.c_int, [_p, _p, _pf64]),
    "sla_norm2": (C.c_int, [_p, _p, _pf64]),
    "sla_vec_add": (C.c_int, [_p, _p, _p, _p]),
    "sla_vec_sub": (C.c_int, [_p, _p, _p, _p]),
    "sla_vec_scale": (C.c_int, [_p, _f64, _p, _p]),
    "sla_vec_axpy": (C.c_int, [_p, _f64, _p, _p, _p]),
    "sla_vec_normalize2": (C.c_int, [_p, _p, _p]),
    "sla_spmv_host": (C.c_int, [_p, _p, _pf64, _pf64]),
    "sla_bicgstab_init": (C.c_int, [_p, _p, _p, _p, _pp]),
    "sla_bicgstab_step": (C.c_int, [_p, _p, _p, _p]),
    "sla_cgs_init": (C.c_int, [_p, _p, _p, _p, _pp]),
    "sla_cgs_step": (C.c_int, [_p, _p, _p, _p]),
    "sla_cgne_init": (C.c_int, [_p, _p, _p, _p, _pp]),
    "sla_cgne_step": (C.c_int, [_p, _p, _p]),
    "sla_krylov_get": (C.c_int, [_p, _p, C.c_int, _pf64]),
    "sla_krylov_view": (C.c_int, [_p, _p, C.c_int, _pp]),
    "sla_krylov_clone": (C.c_int, [_p, _p, _pp]),
    "sla_krylov_free": (None, [_p]),
    "sla_solve_opts_default": (None, [_popts]),
    "sla_linsolve0": (C.c_int, [_p, C.c_int, _p, _p, _p, _popts, _p, _pint, _pf64]),
    "sla_linsolve0_host": (C.c_int, [_p, C.c_int, _p, _pf64, _pf64, _popts, _pf64, _pint, _pf64]),
    "sla_arnoldi": (C.c_int, [_p, _p, _p, C.c_int, _pp, _pf64, _pint]),
    "sla_gmres": (C.c_int, [_p, _p, _p, _p, C.c_int, _popts, _p, _pint, _pf64]),
    "sla_dense_create": (C.c_int, [_p, _i64, _i64, C.c_int, _pp]),
    "sla_dense_from_host": (C.c_int, [_p, _i64, _i64, _pf64, C.c_int, _pp]),
    "sla_dense_generate": (C.c_int, [_p, _i64, _i64, C.c_uint64, C.c_int, _pp]),
    "sla_dense_to_host_f64": (C.c_int, [_p, _p, _pf64]),
    "sla_spmm_dense": (C.c_int, [_p, _p, _p, _p]),
    "sla_spmm_dense_abt": (C.c_int, [_p, _p, _p, _p]),
    "sla_spmm_dense_atb": (C.c_int, [_p, _p, _p, _p]),
    "sla_csr_norm_frobenius": (C.c_int, [_p, _p, _pf64]),
    "sla_dense_column": (C.c_int, [_p, _p, _i64, _pp]),
    "sla_csr_diag_partitions": (C.c_int, [_p, _p, _pp, _pp, _pp]),
    "sla_jacobi_pre": (C.c_int, [_p, _p, _pp]),
    "sla_mssor_pre": (C.c_int, [_p, _p, _f64, _pp, _pp]),
    "sla_ilu0_pre": (C.c_int, [_p, _p, _pp, _pp]),
    "sla_tri_lower_solve": (C.c_int, [_p, _p, _p, _p]),
    "sla_tri_upper_solve": (C.c_int, [_p, _p, _p, _p]),
    "sla_tri_analysis": (C.c_int, [_p, _p, C.c_int, _pint, _pi64]),
    "sla_dense_dims": (C.c_int, [_p, _pi64, _pi64]),
    "sla_dense_to_host": (C.c_int, [_p, _p, _pf64]),
    "sla_dense_free": (None, [_p]),
    # one host process, several GPUs (csrc/multi.cu)
    "sla_init_multi": (C.c_int, [C.c_int, _pint, _pp]),
    "sla_finalize_multi": (None, [_p]),
    "sla_multi_last_error": (C.c_char_p, [_p]),
    "sla_multi_world": (C.c_int, [_p]),
    "sla_multi_ctx": (_p, [_p, C.c_int]),
    "sla_multi_csr_generate": (C.c_int, [_p, C.c_int, _i64, C.c_int, C.c_uint64, _i64, _pp]),
    "sla_multi_csr_from_csr": (C.c_int, [_p, _i64, _i64, _i64, _pi32, _pi32, _pf64, _pp]),
    "sla_multi_csr_dims": (C.c_int, [_p, _pi64, _pi64, _pi64]),
    "sla_multi_csr_free": (None, [_p]),
    "sla_multi_vec_create": (C.c_int, [_p, _i64, _pp]),
    "sla_multi_vec_from_host": (C.c_int, [_p, _i64, _pf64, _pp]),
    "sla_multi_vec_generate": (C.c_int, [_p, _i64, C.c_uint64, _pp]),
    "sla_multi_vec_to_host": (C.c_int, [_p, _p, _pf64]),
    "sla_multi_vec_copy": (C.c_int, [_p, _p, _p]),
    "sla_multi_vec_dim": (_i64, [_p]),
    "sla_multi_vec_free": (None, [_p]),
    "sla_multi_spmv": (C.c_int, [_p, _p, _p, _p]),
    "sla_multi_dot": (C.c_int, [_p, _p, _p, _pf64]),
    "sla_multi_norm2": (C.c_int, [_p, _p, _pf64]),
    "sla_multi_vec_axpy": (C.c_int, [_p, _f64, _p, _p, _p]),
    "sla_multi_vec_scale": (C.c_int, [_p, _f64, _p, _p]),
    "sla_multi_bicgstab_init": (C.c_int, [_p, _p, _p, _p, _pp]),
    "sla_multi_bicgstab_step": (C.c_int, [_p, _p, _p, _p]),
    "sla_multi_cgs_init": (C.c_int, [_p, _p, _p, _p, _pp]),
    "sla_multi_cgs_step": (C.c_int, [_p, _p, _p, _p]),
    "sla_multi_krylov_clone": (C.c_int, [_p, _p, _pp]),
    "sla_multi_krylov_get": (C.c_int, [_p, _p, C.c_int, _pf64]),
    "sla_multi_krylov_free": (None, [_p]),
    "sla_multi_linsolve0": (C.c_int, [_p, C.c_int, _p, _p, _p, _popts, _p, _pint, _pf64]),
    "sla_multi_gmres": (C.c_int, [_p, _p, _p, _p, C.c_int, _popts, _p, _pint, _pf64]),
    "sla_multi_arnoldi": (C.c_int, [_p, _p, _p, C.c_int, _pp, _pf64, _pint]),
    "sla_multi_dense_to_host": (C.c_int, [_p, _p, _pf64]),
    "sla_multi_dense_free": (None, [_p]),
}


def build(force=False):
    """Compile libsla_b200.so in-tree with nvcc for sm_100a (works without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"])
    subprocess.check_call(["make", "-C", csrc, "-j8"])
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). sparse_linear_algebra_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)          # AttributeError here = the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    _lib = L
    return L
