"""(##) dense right operand, BASELINE config 5 shape: A n x n 32 nnz/row, B n x 128 bf16.  usage: bench_spmm.py KIND N [BAND]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sparse_linear_algebra_b200 as sla
kind, n = sys.argv[1], int(sys.argv[2])
band = int(sys.argv[3]) if len(sys.argv) > 3 else 0
k = 128
ctx = sla.default_context()
gk = {"uniform": sla.GEN_UNIFORM, "banded": sla.GEN_BANDED, "block16": sla.GEN_BLOCK16}[kind]
A = sla.SpMatrix.generate(gk, n, 32, 5, band)
rng = np.random.default_rng(0)
Bh = rng.uniform(-1, 1, (min(n, 1 << 16), k))
B = sla.DenseMatrix.fromHost(np.tile(Bh, (n // Bh.shape[0] + 1, 1))[:n], sla.BF16)
Cm = sla.DenseMatrix.zeros(n, k, sla.BF16)
for _ in range(2): A.matMat(B, out=Cm)
ctx.timer_start()
reps = 5
for _ in range(reps): A.matMat(B, out=Cm)
ms = ctx.timer_stop() / reps
nnz = A.nnz
alg = (2 * nnz + 4 * nnz + 4 * (n + 1)) + 2 * n * k + 2 * n * k          # B_spmm, SURVEY.md §8(d)
print(json.dumps({"kind": kind, "n": n, "band": band, "ms": ms, "algorithmic_GB": alg / 1e9, "gbs": alg / ms / 1e6,
                  "gflops": 2 * nnz * k / ms / 1e6, "gather_GBs": nnz * 2 * k / ms / 1e6}))
