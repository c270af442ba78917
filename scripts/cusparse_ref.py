"""Same-box cuSPARSE comparison (SURVEY.md §8(c) "secondary cross-check"): cusparseSpMV (CSR, fp64, int32
indices) on the SAME synthetic matrices bench.py times, called directly through ctypes (torch only allocates the device
buffers and records the events).  Context number only — nothing of the product path runs through cuSPARSE.

    python scripts/cusparse_ref.py [uniform|banded|cfg4 ...]        -> one JSON object on stdout

Library: whatever libcusparse.so.12 the process resolves (torch's bundled copy first, then /usr/local/cuda)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {"uniform": (0, 10_000_000, 32, 0x5EED0002, 0), "banded": (1, 10_000_000, 32, 0x5EED0002, 65536),
         "cfg4": (0, 4_000_000, 64, 0x5EED0004, 0), "cfg3": (2, 4096 * 4096, 5, 0, 4096), "small": (0, 100_000, 16, 7, 0)}
ALGS = {"default": 0, "csr_alg1": 2, "csr_alg2": 3}


def load_cusparse():
    import torch  # noqa: F401  (its import puts the bundled CUDA libraries on the loader path)

    for name in ("libcusparse.so.12", "/usr/local/cuda/lib64/libcusparse.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    try:
        import nvidia.cusparse

        return C.CDLL(os.path.join(os.path.dirname(nvidia.cusparse.__file__), "lib", "libcusparse.so.12"))
    except Exception:
        return None


def cusparse_spmv_ms(lib, m, n, nnz, rp, ci, va, x, y, alg, reps=20):
    import torch

    h, mat, vx, vy = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()

    def ck(st, what):
        if st != 0:
            raise RuntimeError(f"{what}: cusparse status {st}")

    ck(lib.cusparseCreate(C.byref(h)), "create")
    ck(lib.cusparseSetStream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "stream")
    ck(lib.cusparseCreateCsr(C.byref(mat), C.c_int64(m), C.c_int64(n), C.c_int64(nnz), C.c_void_p(rp.data_ptr()), C.c_void_p(ci.data_ptr()),
                             C.c_void_p(va.data_ptr()), 2, 2, 0, 1), "createCsr")          # INDEX_32I x 2, base 0, CUDA_R_64F
    ck(lib.cusparseCreateDnVec(C.byref(vx), C.c_int64(n), C.c_void_p(x.data_ptr()), 1), "dnvec x")
    ck(lib.cusparseCreateDnVec(C.byref(vy), C.c_int64(m), C.c_void_p(y.data_ptr()), 1), "dnvec y")
    alpha, beta = C.c_double(1.0), C.c_double(0.0)
    bsz = C.c_size_t(0)
    ck(lib.cusparseSpMV_bufferSize(h, 0, C.byref(alpha), mat, vx, C.byref(beta), vy, 1, alg, C.byref(bsz)), "bufferSize")
    buf = torch.empty(max(int(bsz.value), 8), dtype=torch.uint8, device="cuda")
    if hasattr(lib, "cusparseSpMV_preprocess"):
        lib.cusparseSpMV_preprocess(h, 0, C.byref(alpha), mat, vx, C.byref(beta), vy, 1, alg, C.c_void_p(buf.data_ptr()))

    def run():
        ck(lib.cusparseSpMV(h, 0, C.byref(alpha), mat, vx, C.byref(beta), vy, 1, alg, C.c_void_p(buf.data_ptr())), "SpMV")

    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.cusparseDestroySpMat(mat); lib.cusparseDestroyDnVec(vx); lib.cusparseDestroyDnVec(vy); lib.cusparseDestroy(h)
    return ms, int(bsz.value)


def measure(case, reps=20):
    """Returns {alg: {ms, gbs}} for one synthetic case, plus this library's own time on the same matrix and the max
    difference of the two results (cuSPARSE sums in another order: a tolerance check, not parity)."""
    import torch

    import sparse_linear_algebra_b200 as sla

    kind, n, k, seed, band = CASES[case]
    A = sla.SpMatrix.generate(kind, n, k, seed, band)
    x = sla.SpVector.generate(n, seed + 1)
    y = sla.SpVector.zeroSV(n)
    ctx = A.ctx
    for _ in range(5):
        A.matVec(x, out=y)
    ctx.timer_start()
    for _ in range(reps):
        A.matVec(x, out=y)
    ours_ms = ctx.timer_stop() / reps
    nbytes = A.spmv_bytes
    rp_h, ci_h, va_h = A.toCSR()
    x_h, y_h = x.toDenseListSV(), y.toDenseListSV()
    nnz = int(ci_h.size)
    del A
    out = {"n": n, "nnz": nnz, "algorithmic_bytes": nbytes, "sla_b200": {"ms": ours_ms, "gbs": nbytes / ours_ms / 1e6}}
    lib = load_cusparse()
    if lib is None:
        out["error"] = "libcusparse.so.12 not loadable"
        return out
    rp, ci, va = torch.from_numpy(rp_h).cuda(), torch.from_numpy(ci_h).cuda(), torch.from_numpy(va_h).cuda()
    xt, yt = torch.from_numpy(x_h).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
    for name, alg in ALGS.items():
        try:
            ms, bsz = cusparse_spmv_ms(lib, n, n, nnz, rp, ci, va, xt, yt, alg, reps)
            diff = float(np.abs(yt.cpu().numpy() - y_h).max() / max(np.abs(y_h).max(), 1e-300))
            out[f"cusparse_{name}"] = {"ms": ms, "gbs": nbytes / ms / 1e6, "workspace_bytes": bsz, "max_rel_diff_vs_sla": diff}
        except Exception as e:
            out[f"cusparse_{name}"] = {"error": str(e)[:200]}
    return out


if __name__ == "__main__":
    cases = sys.argv[1:] or ["uniform", "banded"]
    res = {}
    for cs in cases:
        try:
            res[cs] = measure(cs)
        except Exception as e:
            res[cs] = {"error": str(e)[:300]}
    print(json.dumps(res))
