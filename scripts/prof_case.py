"""Run a handful of SpMV launches on one synthetic case (for ncu captures).
usage: prof_case.py KIND N K BAND [REPS]   KIND in uniform|banded|laplace"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla
kind, n, k, band = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
gk = {"uniform": sla.GEN_UNIFORM, "banded": sla.GEN_BANDED, "laplace": sla.GEN_LAPLACE2D}[kind]
ctx = sla.default_context()
A = sla.SpMatrix.generate(gk, n, k, 2, band)
x = sla.SpVector.generate(n, 3); y = sla.SpVector.zeroSV(n)
for _ in range(reps): A.matVec(x, out=y)
ctx.timer_start()
for _ in range(reps): A.matVec(x, out=y)
ms = ctx.timer_stop() / reps
print(f"{kind} n={n} k={k} band={band}: {ms:.4f} ms  {A.spmv_bytes / ms / 1e6:.1f} GB/s")
