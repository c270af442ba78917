"""Generates tests/golden/e05r0000.npz from the reference's real-matrix fixture.

Run in the authoring container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_e05r0000_golden.py
Inputs : /root/reference/test/data/e05r0000.mtx (236 x 236, 5856 entries, general real coordinate) and
         e05r0000_rhs1.mtx, marshalled exactly as test/Perf.hs:20-45 does (1-based -> 0-based; rhs entries with
         |x| <= 1e-12 dropped by `isNz`; x0 = 0.1).
Outputs: the (i, j, v) triples in file order, the rhs, and what the CPU oracle (the restatement of the reference's
         algorithm) computes from them: CSR arrays, transpose, A #> x, x <# A, and the linSolve0 trajectories.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from sparse_linear_algebra_b200.mmio import read_array, read_matrix_market  # noqa: E402

REF = "/root/reference/test/data"
m, n, i, j, v = read_matrix_market(os.path.join(REF, "e05r0000.mtx"))
rhs = read_array(os.path.join(REF, "e05r0000_rhs1.mtx")).reshape(-1)
A = o.SpMatrix.fromCOO((m, n), i, j, v)
b = o.SpVector.fromListSV(n, [(k, x) for k, x in enumerate(rhs.tolist()) if abs(x) > 1e-12])
x = o.SpVector.synth(0x5EED0E05, n)
rp, ci, va = A.toCSR()
trp, tci, tva = A.transpose().toCSR()
out = dict(m=m, n=n, i=i, j=j, v=v, rhs=rhs, x=x.toDenseListSV(), row_ptr=rp, col=ci, val=va, t_row_ptr=trp, t_col=tci, t_val=tva,
           y=A.matVec(x).toDenseListSV(), yt=A.vecMat(x).toDenseListSV(), dot_xy=np.array([x.dot(A.matVec(x))]))
x0 = o.SpVector.mkSpVR(n, [0.1] * n)
for name, meth in (("bicgstab", o.BICGSTAB_), ("cgs", o.CGS_), ("cgne", o.CGNE_)):
    xs, it, hist = o.linSolve0(meth, A, b, x0, info=True)
    out[f"{name}_iters"] = np.array([it])
    out[f"{name}_hist"] = hist
    out[f"{name}_x"] = xs.toDenseListSV()
st = o.bicgsInit(A, b, x0)
rhat = b - A.matVec(x0)
for k in range(3):
    st = o.bicgstabStep(A, rhat, st)
    out[f"bicgstab_step{k}_x"] = st.x.toDenseListSV()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "e05r0000.npz"), **out)
print({k: (a.shape if hasattr(a, "shape") else a) for k, a in out.items() if k.endswith("iters") or k in ("m", "n")},
      "bicgstab final res", out["bicgstab_hist"][-1] if len(out["bicgstab_hist"]) else None)
