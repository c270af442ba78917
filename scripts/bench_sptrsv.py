"""Times the triangular sweeps (analysis + solve) on the cfg3 Laplacian and the cfg2 uniform matrix.  GPU only.
usage: bench_sptrsv.py [cfg3|cfg2|banded ...]   (default: all three)"""
import json
import sys
import time

sys.path.insert(0, ".")
import sparse_linear_algebra_b200 as sla

ctx = sla.default_context()
out = {}
want = set(sys.argv[1:])
for tag, kind, n, k, band in (("cfg3", sla.GEN_LAPLACE2D, 4096 * 4096, 5, 4096), ("cfg2", sla.GEN_UNIFORM, 10_000_000, 32, 0),
                              ("banded", sla.GEN_BANDED, 10_000_000, 32, 65536)):
    if want and tag not in want:
        continue
    A = sla.SpMatrix.generate(kind, n, k, 0x5EED0002, band)
    b = sla.SpVector.generate(n, 7)
    w = sla.SpVector.zeroSV(n)
    for upper in (False, True):
        t0 = time.perf_counter()
        lv, nzt = A.triAnalysis(upper)
        ctx.sync()
        ta = (time.perf_counter() - t0) * 1e3
        f = sla.triUpperSolve if upper else sla.triLowerSolve
        for _ in range(2):
            f(A, b, out=w)
        ctx.timer_start()
        for _ in range(5):
            f(A, b, out=w)
        ms = ctx.timer_stop() / 5
        nbytes = 12 * nzt + 4 * (n + 1) + 16 * n
        out[f"{tag}_{'upper' if upper else 'lower'}"] = {"analysis_ms": ta, "levels": lv, "nnz_tri": nzt, "solve_ms": ms,
                                                         "us_per_level": ms * 1e3 / lv, "gbs": nbytes / ms / 1e6}
    del A
print(json.dumps(out, indent=1))
