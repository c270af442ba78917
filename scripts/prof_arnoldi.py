"""One arnoldi(A, b, kn) on the cfg-4 family for ncu captures and quick timings.
usage: prof_arnoldi.py [n=4000000] [k=64] [kn=30] [reps=2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparse_linear_algebra_b200 as sla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 64
kn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = sla.default_context()
A = sla.SpMatrix.generate(sla.GEN_UNIFORM, n, k, 0x5EED0004)
b = sla.SpVector.generate(n, 0x5EED0005)
y = sla.SpVector.zeroSV(n)
for _ in range(3): A.matVec(b, out=y)
ctx.timer_start()
for _ in range(10): A.matVec(b, out=y)
spmv_ms = ctx.timer_stop() / 10
for r in range(reps):
    ctx.sync(); t0 = time.perf_counter()
    Q, H, brk = sla.arnoldi(A, b, kn)
    ctx.sync(); dt = time.perf_counter() - t0
    print(f"arnoldi n={n} k={k} kn={kn}: {dt*1e3:.2f} ms wall, (#>) {spmv_ms:.3f} ms each -> non-SpMV {dt*1e3 - kn*spmv_ms:.2f} ms "
          f"for {(16*n*sum(j+1 for j in range(kn)) + 32*n*kn)/1e9:.1f} GB")
    del Q
