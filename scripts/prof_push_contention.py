"""ONE GPU: what does a mode-5 push cost the (#>) panel kernels it runs beside?  (the 0.10 ms still exposed at 8 ranks.)
The push kernels of csrc/p2p.cu copy `mb` MB inside this GPU's memory from a high-priority side stream while the cfg-2 (#>) of an
8-rank row block (1.25M x 10M, 32 nnz/row, uniform; rotated panels 1,1,2,4 through the test hook) runs on the ctx stream.  No NVLink
in the picture: what shows is the SM / shared-memory / HBM side of the contention.
usage: prof_push_contention.py [ROWS=1250000] [MB=70]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 70
n, k, world = 10_000_000, 32, 8
ctx = sla.default_context()
L = ctx.lib
# the row block rank 3 of 8 holds of cfg 2, as an ordinary (non-partitioned) rows x n matrix: same panels, same gathers, no exchange
RANK = 3
hA = C.c_void_p()
ctx.check(L.sla_csr_generate_rows(ctx.h, sla.GEN_UNIFORM, n, k, 0x5EED0002, 0, RANK * rows, (RANK + 1) * rows, C.byref(hA)))
A = sla.SpMatrix(ctx, hA)
x = sla.SpVector.generate(n, 3)
y = sla.SpVector.zeroSV(rows)
src = sla.SpVector.generate(mb * 125_000, 5)
dst = sla.SpVector.zeroSV(mb * 125_000)
h = C.c_void_p()
ctx.check(L.sla_debug_push_create(ctx.h, dst.h, src.h, C.byref(h)))
REPS = 20


def timed(fn):
    for _ in range(3):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(REPS):
        fn()
    return ctx.timer_stop() / REPS


def spmv():
    A.matVec(x, out=y)


def push(ctas, kind):
    def f():
        ctx.check(L.sla_debug_push_start(ctx.h, h, ctas, kind))
        ctx.check(L.sla_debug_push_join(ctx.h, h))
    return f


def both(ctas, kind):
    def f():
        ctx.check(L.sla_debug_push_start(ctx.h, h, ctas, kind))
        A.matVec(x, out=y)
        ctx.check(L.sla_debug_push_join(ctx.h, h))
    return f


out = {"rows": rows, "push_mb": mb, "plans": {}}
for plan in ("auto", "rot1124"):
    if plan == "rot1124":
        ctx.check(L.sla_csr_debug_rot_panels(ctx.h, A.h, world, RANK, None))
    rec = {"panels": L.sla_csr_npanels(A.h), "spmv_alone_ms": timed(spmv), "cases": []}
    for kind, name in ((0, "tma"), (1, "lsu")):
        for ctas in (16, 32, 64, 128, 296):
            p = timed(push(ctas, kind))
            b = timed(both(ctas, kind))
            rec["cases"].append({"kernel": name, "ctas": ctas, "push_alone_ms": round(p, 4), "push_alone_gbs": round(mb * 1e-3 / (p * 1e-3), 1),
                                 "both_ms": round(b, 4), "over_max_ms": round(b - max(p, rec["spmv_alone_ms"]), 4)})
    out["plans"][plan] = rec
    print(json.dumps({plan: rec}), flush=True)
L.sla_debug_push_free(h)
