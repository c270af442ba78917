"""Experiment: uniform-column SpMV GB/s vs n (x footprint vs L2) and vs L2 cache-hint mode."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla
ctx = sla.default_context()
res = []
for n in (1_000_000, 2_000_000, 4_000_000, 6_000_000, 8_000_000, 10_000_000):
    A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, 32, 2)
    x = sla.SpVector.generate(n, 3); y = sla.SpVector.zeroSV(n)
    for _ in range(5): A.matVec(x, out=y)
    ctx.timer_start()
    for _ in range(50): A.matVec(x, out=y)
    ms = ctx.timer_stop() / 50
    res.append({"n": n, "x_MB": 8 * n / 1e6, "ms": round(ms, 4), "gbs": round(A.spmv_bytes / ms / 1e6, 1), "ns_per_nnz": round(ms * 1e6 / (32 * n), 5)})
    del A, x, y
print(json.dumps({"hints": os.environ.get("SLA_SPMV_HINTS", "3"), "sweep": res}))
