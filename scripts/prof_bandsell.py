"""ONE GPU: the sliced-ELL band plan (SLA_SPMV_BAND=3, csrc/spmv_bandsell.cuh) against the tile kernel on the cfg-2 banded family
(10M x 10M, 32 nnz/row, columns within +-65536) and the 4096^2 5-point stencil: time per (#>), algorithmic GB/s, and the FULL result
compared bit for bit with the tile kernel's.  usage: prof_bandsell.py [quick]   (quick: default R / W only, for ncu)"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = sla.default_context()
REPS = 20
PEAK = 6534.8


def run(kind, n, k, band, form, R=None, W=None):
    os.environ["SLA_SPMV_BAND"] = form
    for key, v in (("SLA_BAND_R", R), ("SLA_BAND_W", W)):
        if v is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = str(v)
    t0 = time.perf_counter()
    A = sla.SpMatrix.generate(kind, n, k, 0x5EED0002, band)
    ctx.sync()
    build_s = time.perf_counter() - t0
    x = sla.SpVector.generate(n, 0x5EED0003)
    y = sla.SpVector.zeroSV(n)
    for _ in range(3):
        A.matVec(x, out=y)
    ctx.sync()
    ctx.timer_start()
    for _ in range(REPS):
        A.matVec(x, out=y)
    ms = ctx.timer_stop() / REPS
    global x_keep
    x_keep = x
    return A, y, ms, build_s


cases = [("banded", sla.GEN_BANDED, 10_000_000, 32, 65536)]
x_keep = None
for name, kind, n, k, band in cases:
    A0, y0, ms0, b0 = run(kind, n, k, band, "0")
    ref = y0.toDenseListSV()
    bytes_ = A0.spmv_bytes
    print(json.dumps({"case": name, "plan": "tile kernel", "ms": round(ms0, 4), "gbs": round(bytes_ / ms0 / 1e6, 1),
                      "frac": round(bytes_ / ms0 / 1e6 / PEAK, 4), "build_s": round(b0, 2)}), flush=True)
    del A0
    A, y, ms, b = run(kind, n, k, band, "3")
    VARIANTS = {1: "768 threads, 3 slices x 4 entries", 2: "1024 threads, 2 x 4", 3: "1024 threads, 2 x 3", 4: "512 threads, 4 x 4",
                5: "1024 threads, 2 x 3, next batch prefetched to L2", 6: "1024 threads, 2 x 2, prefetch", 7: "768 threads, 3 x 3, prefetch"}
    for var in ([int(os.environ.get("SLA_BSELL_VARIANT", "3"))] if quick else [3, 5, 6, 7, 2]):
        ctx.set_option("bsell_variant", var)
        y = sla.SpVector.zeroSV(n)                         # fresh output: a variant that wrote nothing must not inherit the previous result
        for _ in range(3):
            A.matVec(x_keep, out=y)
        ctx.sync()
        ctx.timer_start()
        for _ in range(REPS):
            A.matVec(x_keep, out=y)
        ms = ctx.timer_stop() / REPS
        same = y.toDenseListSV().tobytes() == ref.tobytes()
        print(json.dumps({"case": name, "plan": "sliced-ELL band", "variant": var, "shape": VARIANTS[var], "R": 8192, "W": 8192, "ms": round(ms, 4),
                          "gbs": round(bytes_ / ms / 1e6, 1), "frac": round(bytes_ / ms / 1e6 / PEAK, 4), "build_s": round(b, 2),
                          "bit_identical_to_tile_kernel": bool(same)}), flush=True)
    del A, y
