#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: fp64 CSR SpMV GB/s (% of the HBM roofline) and
BiCGSTAB iterations/s, on the configurations BASELINE.json names.

    python bench.py --gpus 1 --steps 200 --warmup 20            # this framework (CUDA, through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 5 --warmup 3   # the reference's CPU algorithm (oracle port)

A step = one (#>) over the synthetic 10M x 10M, 32 nnz/row matrix of SURVEY.md §8(d) config 2 (uniform
columns).  `value` = algorithmic bytes (12 nnz + 20 n + 4) x steps / device time, inputs resident in HBM;
`e2e` = the same metric through the host-buffer C-ABI call sla_spmv_host (pinned host x -> device, kernel,
device y -> host inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "csr_spmv_fp64_gbs"
N_CFG2, K_CFG2, SEED_CFG2 = 10_000_000, 32, 0x5EED0002
G_CFG3 = 4096
FALLBACK_HBM_GBS = 6650.0


def load_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def load_traffic():
    """DRAM bytes of the cfg-2 (#>) from the committed ncu --set full capture (profiles/traffic.json): the step is TWO
    launches of spmv_tile_kernel (one per column panel), so the per-step figure — the one comparable with
    `algorithmic_bytes_per_step` — is the sum over both.  Returns (per_step, per_launch, launches) or Nones."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        per_launch = d.get("spmv_cfg2_dram_bytes_per_launch")
        launches = d.get("spmv_cfg2_launches_per_step", 2)
        return d.get("spmv_cfg2_dram_bytes_per_step", per_launch * launches if per_launch else None), per_launch, launches
    except Exception:
        return None, None, None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def spmv_bytes(n, nnz):
    return 12 * nnz + 20 * n + 4


def cpu_baseline(threads, n_sample=2_000_000, reps=3):
    """The oracle's (#>) (per-row ordered intersection + left fold, the reference's algorithm) timed on the
    host cores on a bounded sample of config 2: same family, n_sample rows x 32 nnz/row."""
    from oracle import oracle as ora

    ora.build()
    A = ora.SpMatrix.synth(ora.GEN_UNIFORM, n_sample, K_CFG2, SEED_CFG2)
    x = ora.SpVector.synth(SEED_CFG2 + 1, n_sample)
    ora.time_matvec(A, x, 1, threads)
    sec = ora.time_matvec(A, x, reps, threads)
    gbs = spmv_bytes(n_sample, n_sample * K_CFG2) / sec / 1e9
    # the reference itself is single-threaded (cabal: no -threaded): the same product on ONE host thread, for context
    sec1 = ora.time_matvec(A, x, 1, 1)
    gbs1 = spmv_bytes(n_sample, n_sample * K_CFG2) / sec1 / 1e9
    return gbs, sec, f"config-2 family (uniform, 32 nnz/row) at n={n_sample} rows ({n_sample * K_CFG2} nnz), {reps} matvecs", gbs1


def scipy_sanity(n_sample=500_000):
    """Third leg of BASELINE.md §4: scipy's CSR product (one thread) on the same family — an implementation that shares
    no code with the oracle.  Returns GB/s of algorithmic bytes, or None when scipy is unavailable."""
    try:
        import scipy.sparse as sp

        from oracle import oracle as ora

        A = ora.SpMatrix.synth(ora.GEN_UNIFORM, n_sample, K_CFG2, SEED_CFG2)
        rp, ci, va = A.toCSR()
        M = sp.csr_matrix((np.asarray(va), np.asarray(ci), np.asarray(rp)), shape=(n_sample, n_sample))
        x = np.asarray(ora.SpVector.synth(SEED_CFG2 + 1, n_sample).toDenseListSV())
        M @ x
        t0 = time.perf_counter()
        for _ in range(3):
            M @ x
        sec = (time.perf_counter() - t0) / 3
        return spmv_bytes(n_sample, n_sample * K_CFG2) / sec / 1e9
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path (oracle port; the Haskell cannot be
    built here: no GHC), all host threads, each step a bounded sample of config 2."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as ora

    ora.build()
    threads = os.cpu_count() or 1
    n_sample = 2_000_000
    A = ora.SpMatrix.synth(ora.GEN_UNIFORM, n_sample, K_CFG2, SEED_CFG2)
    x = ora.SpVector.synth(SEED_CFG2 + 1, n_sample)
    for _ in range(args.warmup):
        ora.time_matvec(A, x, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.time_matvec(A, x, 1, threads)
    sec = (time.perf_counter() - t0) / args.steps
    gbs = spmv_bytes(n_sample, n_sample * K_CFG2) / sec / 1e9
    sample = f"config-2 family (uniform, 32 nnz/row) at n={n_sample} rows per step; GB/s is size-normalised"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: CSR SpMV (#>) fp64, 10M x 10M, 32 nnz/row, uniform columns (bounded sample)",
                   "sample_rows": n_sample, "nnz_per_row": K_CFG2},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sparse_linear_algebra_b200 as sla
    from sparse_linear_algebra_b200 import dist as sd

    if world > 1:
        ctx = sd.init_context(local)
    else:
        ctx = sla.Context(local)
        sla.set_default_context(ctx)

    def barrier():
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()
        ctx.sync()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gen(kind, n, k, seed, band=0):
        if world == 1:
            A = sla.SpMatrix.generate(kind, n, k, seed, band)
            A.row_starts = [0, n]
            return A
        return sd.generate_distributed(ctx, kind, n, k, seed, band)

    def vec(n, seed, starts):
        return sd.generate_vector_slice(ctx, n, seed, starts, rank)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = ctx.launches
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()
        barrier()
        return max_over_ranks(ms) / steps, ctx.launches - l0

    # ---- config 2: the SAME 10M x 10M matrix at every N (strong scaling), row-partitioned over the ranks;
    # each (#>) includes the exchange of the x entries the local rows reference.
    n, k = N_CFG2, K_CFG2
    nbytes = spmv_bytes(n, n * k)                         # algorithmic bytes of the GLOBAL product
    A = gen(sla.GEN_UNIFORM, n, k, SEED_CFG2)
    starts = A.row_starts
    nloc = starts[rank + 1] - starts[rank]
    x = vec(n, SEED_CFG2 + 1, starts)
    y = sla.SpVector.zeroSV(nloc)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step, launches = timed(lambda: A.matVec(x, out=y), args.steps, args.warmup)
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    kernel_gbs = value / world                             # per-GPU share of the algorithmic bytes

    # ---- e2e: host buffers through sla_spmv_host (pinned x slice -> device, exchange + kernel, y slice -> host)
    import ctypes as C

    xh = ctx.pinned(max(nloc, 1))
    yh = ctx.pinned(max(nloc, 1))
    xh[:nloc] = x.toDenseListSV()
    pd = C.POINTER(C.c_double)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_call():
        ctx.check(ctx.lib.sla_spmv_host(ctx.h, A.h, xh.ctypes.data_as(pd), yh.ctypes.data_as(pd)))

    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_gbs = nbytes / e2e_s / 1e9
    clocks = sampler.stop() if rank == 0 else None
    exchange = getattr(A, "dist_plan", [])
    recv_bytes = 8 * sum(c for d, _, _, c in exchange if d == 0)

    extra = {"x_exchange_recv_bytes_per_rank": recv_bytes}
    if world > 1:
        extra["collectives"] = {"allreduce": "peer-memory kernel over NVLink (csrc/p2p.cu)" if getattr(ctx, "p2p", False) else "nccl",
                                "x_exchange": {1: "peer-memory push kernel (csrc/p2p.cu)",
                                               2: "copy-engine all-gather consumed in arrival order (csrc/p2p.cu mode 2)"}.get(
                                    getattr(A, "dist_p2p_mode", 0), "ncclAllGather" if getattr(A, "dist_allgather", False) else "nccl send/recv")}

    def sptrsv_extra(tag, M, rhs):
        if args.no_sptrsv:
            return
        try:
            _sptrsv_extra(tag, M, rhs)
        except Exception as e:                     # context numbers only: never lose the headline line
            extra[f"sptrsv_{tag}_error"] = str(e)[:200]

    def _sptrsv_extra(tag, M, rhs):
        """Triangular sweep (triLowerSolve, Sparse.hs:750-777): latency-bound by the number of dependency levels
        (8191 wavefronts on the 4096^2 grid); single GPU."""
        t0 = time.perf_counter()
        lv, nzt = M.triAnalysis(False)
        ctx.sync()
        extra[f"sptrsv_{tag}_analysis_ms"] = (time.perf_counter() - t0) * 1e3
        wv = sla.SpVector.zeroSV(rhs.dim)
        mst, _ = timed(lambda: sla.triLowerSolve(M, rhs, out=wv), 5, 2)
        extra[f"sptrsv_{tag}_lower_ms"] = mst
        extra[f"sptrsv_{tag}_levels"] = lv
        extra[f"sptrsv_{tag}_us_per_level"] = mst * 1e3 / max(lv, 1)
        extra[f"sptrsv_{tag}_gbs"] = (12 * nzt + 4 * (rhs.dim + 1) + 16 * rhs.dim) / (mst * 1e-3) / 1e9

    want = set() if args.quick else set(args.extras.split(","))
    if args.no_sptrsv:
        want.discard("sptrsv")
    if "sptrsv" in want and world == 1:
        sptrsv_extra("cfg2", A, x)
    del A
    if "banded" in want:
        # ---- banded variant of config 2 (columns within +-65536 of the row)
        B = gen(sla.GEN_BANDED, n, k, SEED_CFG2, 65536)
        msb, _ = timed(lambda: B.matVec(x, out=y), args.steps, args.warmup)
        extra["spmv_banded_gbs"] = nbytes / (msb * 1e-3) / 1e9
        extra["spmv_banded_ms"] = msb
        del B
    if "cfg3" in want:
        # ---- config 3: BiCGSTAB on the 5-point Laplacian 4096^2, fixed number of bicgstabStep calls
        g = G_CFG3
        n3 = g * g
        L3 = gen(sla.GEN_LAPLACE2D, n3, 5, 0, g)
        s3 = L3.row_starts
        n3loc = s3[rank + 1] - s3[rank]
        xt = vec(n3, 3, s3)
        b = L3 @ xt
        st = sla.bicgsInit(L3, b, sla.SpVector.zeroSV(n3loc))
        rhat = st.r.copy()
        its = max(10, min(args.steps, 100))
        ms3, l3 = timed(lambda: sla.bicgstabStep(L3, rhat, st), its, 5)
        nnz3 = 5 * n3 - 4 * g
        b3 = 24 * nnz3 + 168 * n3                          # B_bicgstab_step, SURVEY.md §8(d)
        extra["bicgstab_cfg3_iters_per_s"] = 1e3 / ms3
        extra["bicgstab_cfg3_ms_per_iter"] = ms3
        extra["bicgstab_cfg3_gbs"] = b3 / (ms3 * 1e-3) / 1e9
        extra["bicgstab_cfg3_launches_per_iter"] = l3 / its
        ms3s, _ = timed(lambda: L3.matVec(xt, out=b), its, 3)
        extra["spmv_cfg3_gbs"] = spmv_bytes(n3, nnz3) / (ms3s * 1e-3) / 1e9
        if world == 1 and "sptrsv" in want:
            sptrsv_extra("cfg3", L3, b)
        del L3, st
    if "cfg4" in want:
        # ---- config 4: arnoldi(A, b, 30) on the random non-symmetric 4M x 4M, 64 nnz/row matrix (row-partitioned)
        n4, k4 = 4_000_000, 64
        A4 = gen(sla.GEN_UNIFORM, n4, k4, 0x5EED0004)
        b4 = vec(n4, 0x5EED0005, A4.row_starts)
        sla.arnoldi(A4, b4, 4)                              # warm-up
        s4 = float("inf")
        for _ in range(2):                                  # best of two cycles: the call allocates its 1 GB basis (cudaMalloc jitter)
            Qd = None
            barrier()
            t0 = time.perf_counter()
            Qd, H4, brk4 = sla.arnoldi(A4, b4, 30)
            barrier()
            s4 = min(s4, max_over_ranks(time.perf_counter() - t0))
        b4bytes = 30 * spmv_bytes(n4, n4 * k4) + 8400 * n4    # B_arnoldi cycle, SURVEY.md §8(d)
        extra["arnoldi_cfg4_steps_per_s"] = 30 / s4
        extra["arnoldi_cfg4_ms_per_cycle"] = s4 * 1e3
        extra["arnoldi_cfg4_gbs"] = b4bytes / s4 / 1e9
        del A4, Qd
    if "cfg5" in want:
        # ---- config 5: (##) CSR 10M x 10M x dense 10M x 128 bf16 (single GPU here; the row-partitioned (##) exists but has not
        # been run on hardware yet, so it stays out of the default multi-GPU line).
        # K16 = block-structured family (tcgen05 tile path); U = uniform columns (gather kernel, L2-bound).
        if world == 1:
            k5 = 128
            B5 = sla.DenseMatrix.generate(n, k5, 0x5EED0055, sla.BF16)
            C5 = sla.DenseMatrix.zeros(n, k5, sla.BF16)
            for tag, kind in (("k16", sla.GEN_BLOCK16), ("uniform", sla.GEN_UNIFORM)):
                A5 = sla.SpMatrix.generate(kind, n, 32, 0x5EED0005)
                reps = 5 if tag == "k16" else 2
                ms5, _ = timed(lambda: A5.matMat(B5, out=C5), reps, 1)
                b5 = 6 * A5.nnz + 4 * (n + 1) + 4 * n * k5        # B_spmm, SURVEY.md §8(d): bf16 A values, B and C once
                extra[f"spmm_cfg5_{tag}_ms"] = ms5
                extra[f"spmm_cfg5_{tag}_gbs"] = b5 / (ms5 * 1e-3) / 1e9
                extra[f"spmm_cfg5_{tag}_tflops"] = 2 * A5.nnz * k5 / (ms5 * 1e-3) / 1e12
                del A5
            del B5, C5
        elif os.environ.get("SLA_BENCH_DIST_SPMM") == "1":
            # row-partitioned (##): written after the round-1 GPU budget was spent, so it is opt-in and can never cost the line
            try:
                k5 = 128
                A5 = gen(sla.GEN_BLOCK16, n, 32, 0x5EED0005)
                s5 = A5.row_starts
                m5 = s5[rank + 1] - s5[rank]
                B5 = sla.DenseMatrix.generate(m5, k5, 0x5EED0055 + rank, sla.BF16)      # this rank's row slice of B
                C5 = sla.DenseMatrix.zeros(m5, k5, sla.BF16)
                ms5, _ = timed(lambda: A5.matMat(B5, out=C5), 5, 1)
                b5 = 6 * n * 32 + 4 * (n + 1) + 4 * n * k5
                extra["spmm_cfg5_k16_dist_ms"] = ms5
                extra["spmm_cfg5_k16_dist_gbs"] = b5 / (ms5 * 1e-3) / 1e9
                del A5, B5, C5
            except Exception as e:
                extra["spmm_cfg5_dist_error"] = str(e)[:200]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peak()
    for key in ("spmv_banded_gbs", "bicgstab_cfg3_gbs", "spmv_cfg3_gbs", "arnoldi_cfg4_gbs", "spmm_cfg5_k16_gbs", "spmm_cfg5_uniform_gbs",
                "sptrsv_cfg3_gbs", "sptrsv_cfg2_gbs"):
        if key in extra:
            extra[key.replace("_gbs", "_frac_per_gpu")] = extra[key] / world / peak
    cpu = None
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        gbs, sec, sample, gbs1 = cpu_baseline(threads)
        cpu = {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample,
               "seconds_per_matvec_on_sample": sec, "single_thread_value": gbs1, "scipy_csr_single_thread_value": scipy_sanity()}
    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: CSR SpMV (#>) fp64, 10M x 10M, 32 nnz/row, uniform columns"
                               + (f", row-partitioned over {world} B200 (remote x blocks exchanged every step)" if world > 1 else ", 1 x B200"),
                   "n": n, "nnz": n * k, "algorithmic_bytes_per_step": nbytes,
                   "l2": "no flush: the 3.84 GB matrix stream exceeds the 126 MB L2 every step"},
        "roofline": {"bound": "hbm", "achieved": kernel_gbs, "peak": peak, "unit": "GB/s", "frac": kernel_gbs / peak,
                     "traffic": load_traffic()[0] if world == 1 else None,     # the ncu capture is of the single-GPU step
                     "traffic_per_launch": load_traffic()[1] if world == 1 else None, "launches_per_step": load_traffic()[2],
                     "traffic_note": "achieved and traffic are per STEP = one (#>) = launches_per_step launches of the kernel",
                     "peak_source": peak_src, "per_gpu": True,
                     "kernel": "spmv_tile_kernel<1024, EPI_NONE> (one launch per column panel, 2 panels at n = 10M)"},
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": 8 * nloc, "d2h_bytes_per_step": 8 * nloc,
                "ms_per_step": e2e_s * 1e3, "api": "sla_spmv_host (pinned host buffers, per-rank slices)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "cpu_baseline": cpu,
        "extra": extra,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="skip the banded / BiCGSTAB context runs")
    ap.add_argument("--extras", default="sptrsv,banded,cfg3,cfg4,cfg5",
                    help="comma list of the context runs to include (sptrsv, banded, cfg3, cfg4, cfg5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sptrsv", action="store_true", help="skip the triangular-sweep context numbers")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
