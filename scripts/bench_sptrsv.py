"""Times the triangular sweeps (analysis + solve) on the cfg3 Laplacian, the cfg2 uniform matrix and its banded variant.  GPU only.
usage: bench_sptrsv.py [--sweep] [cfg3|cfg2|banded ...]   (default: all three)
--sweep also times the kernel variants (SLA_TRI_MODE / SLA_TRI_LIF / SLA_TRI_BACKOFF are read at every solve)."""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import sparse_linear_algebra_b200 as sla

ctx = sla.default_context()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
sweep = "--sweep" in sys.argv
want = set(args)
VARIANTS = [{}]
if sweep:
    VARIANTS = [{"SLA_TRI_MODE": "0", "SLA_TRI_BACKOFF": "0"}, {"SLA_TRI_MODE": "0", "SLA_TRI_BACKOFF": "100"},
                {"SLA_TRI_MODE": "0", "SLA_TRI_BACKOFF": "500"},
                {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "2", "SLA_TRI_BACKOFF": "0"}, {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "4", "SLA_TRI_BACKOFF": "0"},
                {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "8", "SLA_TRI_BACKOFF": "0"}, {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "16", "SLA_TRI_BACKOFF": "0"},
                {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "4", "SLA_TRI_BACKOFF": "100"}, {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "8", "SLA_TRI_BACKOFF": "100"},
                {"SLA_TRI_MODE": "1", "SLA_TRI_LIF": "64", "SLA_TRI_BACKOFF": "0"}]
out = {}
for tag, kind, n, k, band in (("cfg3", sla.GEN_LAPLACE2D, 4096 * 4096, 5, 4096), ("cfg2", sla.GEN_UNIFORM, 10_000_000, 32, 0),
                              ("banded", sla.GEN_BANDED, 10_000_000, 32, 65536)):
    if want and tag not in want:
        continue
    A = sla.SpMatrix.generate(kind, n, k, 0x5EED0002, band)
    b = sla.SpVector.generate(n, 7)
    w = sla.SpVector.zeroSV(n)
    ref = {}
    for upper in (False, True):
        t0 = time.perf_counter()
        lv, nzt = A.triAnalysis(upper)
        ctx.sync()
        ta = (time.perf_counter() - t0) * 1e3
        f = sla.triUpperSolve if upper else sla.triLowerSolve
        nbytes = 12 * nzt + 4 * (n + 1) + 16 * n
        for var in VARIANTS:
            for key in ("SLA_TRI_MODE", "SLA_TRI_LIF", "SLA_TRI_BACKOFF"):
                os.environ.pop(key, None)
            os.environ.update(var)
            for _ in range(2):
                f(A, b, out=w)
            reps = 3 if sweep else 5
            ctx.timer_start()
            for _ in range(reps):
                f(A, b, out=w)
            ms = ctx.timer_stop() / reps
            got = w.toDenseListSV()
            same = True
            if upper not in ref:
                ref[upper] = got.tobytes()
            else:
                same = ref[upper] == got.tobytes()       # every variant must produce the same bits
            name = f"{tag}_{'upper' if upper else 'lower'}" + ("" if not var else "_" + "_".join(f"{k[8:].lower()}{v}" for k, v in var.items()))
            out[name] = {"analysis_ms": round(ta, 2), "levels": lv, "nnz_tri": nzt, "solve_ms": round(ms, 4),
                         "us_per_level": round(ms * 1e3 / lv, 3), "gbs": round(nbytes / ms / 1e6, 1), "same_bits": same}
    del A
print(json.dumps(out, indent=1))
