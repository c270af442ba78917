"""GMRES(30) fixed work on the cfg-4 family: wall clock vs stream time vs launch count (is the cycle GPU-bound or host-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = sla.default_context()
A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, 0x5EED0004)
b = sla.SpVector.generate(n, 0x5EED0005)
x0 = sla.SpVector.zeroSV(n)
sla.gmres(A, b, x0, restart=30, nits=30, fixed_work=True)
for nits in (30, 60, 300):
    ctx.sync(); l0 = ctx.launches
    ctx.timer_start(); t0 = time.perf_counter()
    x, it, res = sla.gmres(A, b, x0, restart=30, nits=nits, fixed_work=True, info=True)
    ctx.sync(); wall = time.perf_counter() - t0
    ms = ctx.timer_stop()
    print(f"gmres nits={nits}: wall {wall*1e3:.1f} ms, stream {ms:.1f} ms, launches {ctx.launches - l0}, iters {it}, res {res:.3e} -> {wall*1e3/ (nits/30):.1f} ms per cycle")
for _ in range(2):
    ctx.sync(); t0 = time.perf_counter()
    Q, H, brk = sla.arnoldi(A, b, 30)
    ctx.sync(); print(f"arnoldi 30: wall {(time.perf_counter()-t0)*1e3:.1f} ms")
    del Q
