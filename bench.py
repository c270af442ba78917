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


def kernel_source_stamp():
    """sha256[:16] of the (#>) kernel source: profiles/traffic.json is only quoted while it was measured on this source."""
    import hashlib

    try:
        with open(os.path.join(ROOT, "sparse_linear_algebra_b200", "csrc", "spmv.cu"), "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except Exception:
        return None


def load_traffic():
    """DRAM bytes of the cfg-2 (#>) from the committed ncu --set full capture (profiles/traffic.json): the step is TWO
    launches of spmv_tile_kernel (one per column panel), so the per-step figure — the one comparable with
    `algorithmic_bytes_per_step` — is the sum over both.  The file carries the stamp of the kernel source it was captured
    on (`spmv_cu_sha16`); a capture of another source is stale and is NOT quoted.  Returns (per_step, per_launch, launches,
    note)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        launches = d.get("spmv_cfg2_launches_per_step", 2)
        if d.get("spmv_cu_sha16") != kernel_source_stamp():
            return None, None, launches, "profiles/traffic.json was captured on another version of csrc/spmv.cu: stale, not quoted"
        per_launch = d.get("spmv_cfg2_dram_bytes_per_launch")
        return (d.get("spmv_cfg2_dram_bytes_per_step", per_launch * launches if per_launch else None), per_launch, launches,
                f"ncu --set full capture of this kernel source ({d.get('captured', 'undated')})")
    except Exception:
        return None, None, None, "no capture"


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


def host_sample_rows():
    """Rows of the config-2 matrix the CPU legs run on: the full 10M when the host has the memory for the oracle's
    containers (~6 GB, built in ~25 s), else a 2M-row sample of the same family."""
    try:
        with open("/proc/meminfo") as f:
            avail_kb = next(int(line.split()[1]) for line in f if line.startswith("MemAvailable"))
        return N_CFG2 if avail_kb >= 24 * 1024 * 1024 else 2_000_000
    except Exception:
        return 2_000_000


def cpu_baseline(threads, n_sample=None, reps=3):
    """The oracle's (#>) (per-row ordered intersection + left fold, the reference's algorithm) timed on the
    host cores on config 2 itself when the host memory allows (else a bounded sample of the same family)."""
    from oracle import oracle as ora

    ora.build()
    n_sample = n_sample or host_sample_rows()
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
    n_sample = host_sample_rows()
    A = ora.SpMatrix.synth(ora.GEN_UNIFORM, n_sample, K_CFG2, SEED_CFG2)
    x = ora.SpVector.synth(SEED_CFG2 + 1, n_sample)
    for _ in range(args.warmup):
        ora.time_matvec(A, x, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.time_matvec(A, x, 1, threads)
    sec = (time.perf_counter() - t0) / args.steps
    gbs = spmv_bytes(n_sample, n_sample * K_CFG2) / sec / 1e9
    sample = (f"config 2 itself: n={n_sample} rows, {n_sample * K_CFG2} nnz per step" if n_sample == N_CFG2 else
              f"config-2 family (uniform, 32 nnz/row) at n={n_sample} rows per step; GB/s is size-normalised")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: CSR SpMV (#>) fp64, 10M x 10M, 32 nnz/row, uniform columns" + ("" if n_sample == N_CFG2 else " (bounded sample)"),
                   "n": n_sample, "nnz": n_sample * K_CFG2, "sample_rows": n_sample, "nnz_per_row": K_CFG2,
                   "algorithmic_bytes_per_step": spmv_bytes(n_sample, n_sample * K_CFG2)},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def _seq_row_dot(cols, vals, xh):
    """The reference's row product: strict left fold from 0 over ascending columns, products rounded once (Common.hs:259-260)."""
    acc = 0.0
    for cj, v in zip(cols.tolist(), vals.tolist()):
        acc = acc + v * xh[cj]
    return acc


def parity_rows(y_local, row0, nloc, n, k, seed, kind, band, xh, exact, nrows=304):
    """Checks `nrows` rows of this rank's slice of y = A x against rows regenerated by the oracle's generator
    (oracle.synth_row) and folded the reference's way.  exact=True: the bits must match; else the SURVEY.md section 8(d) bound
    |dy| <= (k_i + 2) u sum |a_ij x_j| (arrival-order exchange: same products, rotated fold).  Outside every timed region."""
    from oracle import oracle as ora

    rng = np.random.default_rng(1234 + row0)
    rows = np.unique(np.concatenate([[0, 1, max(nloc - 1, 0), max(nloc - 2, 0)], rng.integers(0, max(nloc, 1), nrows - 4)]))
    bit_exact, worst, bad = 0, 0.0, 0
    for rl in rows.tolist():
        if rl >= nloc:
            continue
        cols, vals = ora.synth_row(kind, n, k, seed, band, row0 + rl)
        ref = _seq_row_dot(cols, vals, xh)
        got = float(y_local[rl])
        if got == ref:
            bit_exact += 1
            continue
        bound = (len(cols) + 2) * 2.0 ** -53 * float(np.sum(np.abs(vals) * np.abs(xh[cols])))
        ratio = abs(got - ref) / bound if bound > 0 else float("inf")
        worst = max(worst, ratio)
        if exact or ratio > 1.0:
            bad += 1
    return {"rows": int(len(rows)), "bit_exact_rows": bit_exact, "max_err_over_bound": worst, "bad_rows": bad}


def parity_cfg2(A, y, starts, rank, world, dist, n, k):
    """Parity of THIS run's cfg-2 result, outside the timed region: sampled rows of every rank's slice of y against rows
    regenerated by the oracle (the checker; nothing of the timed path touches it)."""
    from oracle import oracle as ora

    ora.build()
    nloc = starts[rank + 1] - starts[rank]
    xh_full = np.asarray(ora.SpVector.synth(SEED_CFG2 + 1, n).toDenseListSV())
    p2p_mode = getattr(A, "dist_p2p_mode", 0)
    mine = parity_rows(y.toDenseListSV(), starts[rank], nloc, n, k, SEED_CFG2, ora.GEN_UNIFORM, 0, xh_full, exact=(p2p_mode not in (2, 5)))
    del xh_full
    if dist is not None:
        allp = [None] * world
        dist.all_gather_object(allp, mine)
    else:
        allp = [mine]
    return {"what": "rows of y = A x (cfg 2, this run) vs oracle.synth_row folded left to right",
            "rows": sum(q["rows"] for q in allp), "bit_exact_rows": sum(q["bit_exact_rows"] for q in allp),
            "max_err_over_bound": max(q["max_err_over_bound"] for q in allp),
            "criterion": "bit-exact" if p2p_mode not in (2, 5) else "(k+2) u sum|a_ij x_j| (arrival-order / two-phase exchange folds each row in rotated column order)",
            "ok": all(q["bad_rows"] == 0 for q in allp)}


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

    # ---- per-step spread (each step individually synchronised; the headline is the back-to-back loop above)
    each = []
    for _ in range(min(args.steps, 20)):
        barrier()
        ctx.timer_start()
        A.matVec(x, out=y)
        each.append(max_over_ranks(ctx.timer_stop()))
    step_stats = {"min": float(np.min(each)), "median": float(np.median(each)), "max": float(np.max(each)), "n": len(each),
                  "note": "steps timed one by one with a synchronisation (and at N > 1 a barrier) between them"}

    parity = parity_cfg2(A, y, starts, rank, world, dist, n, k)

    # ---- what the exchange costs: the same kernels WITHOUT the x exchange (diagnostic switch, results invalid)
    xch = None
    if world > 1:
        ctx.set_option("skip_exchange", 1)
        ms_kernels, _ = timed(lambda: A.matVec(x, out=y), max(10, args.steps // 2), 3)
        ctx.set_option("skip_exchange", 0)
        A.matVec(x, out=y)
        xch = {"step_ms": ms_per_step, "kernels_only_ms": ms_kernels, "exposed_exchange_ms": ms_per_step - ms_kernels,
               "mode": getattr(A, "dist_p2p_mode", 0)}

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
                                               2: "copy-engine all-gather consumed in arrival order (csrc/p2p.cu mode 2)",
                                               3: "LL halo kernel (csrc/p2p.cu mode 3)",
                                               4: "copy-engine all-gather waited for as a whole (csrc/p2p.cu mode 4)",
                                               5: "two-phase peer-memory push under rotated near/far panels (csrc/p2p.cu mode 5)"}.get(
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
        if world == 1:
            # the same matrix under the opt-in sliced-ELL band plan (x staged through shared memory, csrc/spmv_bandsell.cuh): timed, and
            # its result compared bit for bit with the tile kernel's on all n rows.  Never allowed to break the line.
            try:
                yb_ref = y.toDenseListSV()
                prev = os.environ.get("SLA_SPMV_BAND")
                os.environ["SLA_SPMV_BAND"] = "3"
                try:
                    t0 = time.perf_counter()
                    Bs = gen(sla.GEN_BANDED, n, k, SEED_CFG2, 65536)
                    ctx.sync()
                    build_s = time.perf_counter() - t0
                finally:
                    if prev is None:
                        os.environ.pop("SLA_SPMV_BAND", None)
                    else:
                        os.environ["SLA_SPMV_BAND"] = prev
                ys = sla.SpVector.zeroSV(n)
                mss, _ = timed(lambda: Bs.matVec(x, out=ys), args.steps, args.warmup)
                extra["spmv_banded_sell"] = {"ms": mss, "gbs": nbytes / (mss * 1e-3) / 1e9, "frac": nbytes / (mss * 1e-3) / 1e9 / load_peak()[0],
                                             "plan_build_s": build_s, "bit_identical_to_tile_kernel": ys.toDenseListSV().tobytes() == yb_ref.tobytes(),
                                             "how": "SLA_SPMV_BAND=3 (opt-in): sliced-ELL cells, x in shared memory, a thread per row"}
                del Bs, ys
            except Exception as exc:                        # noqa: BLE001 - context number only
                extra["spmv_banded_sell"] = {"error": str(exc)[:200]}
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
        # parity of this run: after its+5 steps the recurrence residual still equals the true residual b - A x (size-independent
        # property: it exercises the distributed (#>), both fused dots and the all-reduced scalars of every step)
        r0n = rhat.norm2()
        true_res = ((L3 @ st.x) - b).norm2()
        rec_res = st.r.norm2()
        extra["bicgstab_cfg3_parity"] = {"steps": its + 5, "true_residual": true_res, "recurrence_residual": rec_res, "r0": r0n,
                                         "ok": bool(abs(true_res - rec_res) <= 1e-8 * r0n and np.isfinite(true_res))}
        ms3s, _ = timed(lambda: L3.matVec(xt, out=b), its, 3)
        extra["spmv_cfg3_gbs"] = spmv_bytes(n3, nnz3) / (ms3s * 1e-3) / 1e9
        if world > 1:
            ctx.set_option("skip_exchange", 1)
            ms3k, _ = timed(lambda: L3.matVec(xt, out=b), its, 3)
            ctx.set_option("skip_exchange", 0)
            extra["spmv_cfg3_exchange"] = {"step_ms": ms3s, "kernels_only_ms": ms3k, "exposed_exchange_ms": ms3s - ms3k,
                                           "mode": getattr(L3, "dist_p2p_mode", 0)}
            # a 256^2 Laplacian through the same distributed code path against the oracle's trajectory (5 steps)
            from oracle import oracle as ora

            gs = 256
            Ls = gen(sla.GEN_LAPLACE2D, gs * gs, 5, 0, gs)
            ss = Ls.row_starts
            xs = vec(gs * gs, 3, ss)
            bs = Ls @ xs
            sts = sla.bicgsInit(Ls, bs, sla.SpVector.zeroSV(ss[rank + 1] - ss[rank]))
            rhs = sts.r.copy()
            Lo = ora.SpMatrix.synth(ora.GEN_LAPLACE2D, gs * gs, 5, 0, gs)
            xo_ = ora.SpVector.synth(3, gs * gs)
            bo_ = Lo.matVec(xo_)
            x0o = ora.SpVector.mkSpVR(gs * gs, np.zeros(gs * gs))
            sto = ora.bicgsInit(Lo, bo_, x0o)
            rho_ = bo_ - Lo.matVec(x0o)
            worst = 0.0
            for _ in range(5):
                sla.bicgstabStep(Ls, rhs, sts)
                sto = ora.bicgstabStep(Lo, rho_, sto)
                xr_ = np.asarray(sto.x.toDenseListSV())
                worst = max(worst, float(np.abs(sts.x.toDenseListSV() - xr_[ss[rank]:ss[rank + 1]]).max() / max(np.abs(xr_).max(), 1e-300)))
            extra["bicgstab_small_vs_oracle"] = {"grid": gs, "steps": 5, "max_rel_diff": max_over_ranks(worst), "ok": max_over_ranks(worst) <= 1e-10}
            del Ls, sts
        if world == 1 and "sptrsv" in want:
            sptrsv_extra("cfg3", L3, b)
        del L3, st
    if "cfg4" in want:
        # ---- config 4: arnoldi(A, b, 30) on the random non-symmetric 4M x 4M, 64 nnz/row matrix (row-partitioned)
        n4, k4 = 4_000_000, 64
        A4 = gen(sla.GEN_UNIFORM, n4, k4, 0x5EED0004)
        b4 = vec(n4, 0x5EED0005, A4.row_starts)
        sla.arnoldi(A4, b4, 4)                              # warm-up
        s4, cyc4 = float("inf"), []
        sampler4a = ClockSampler(local)
        if rank == 0:
            sampler4a.start()
        for _ in range(3):                                  # best of three cycles, all reported: the call allocates its 1 GB basis
            Qd = None                                       # (cudaMalloc jitter) and these streaming kernels sit at the board's
            barrier()                                       # power cap, so one cycle can run at a lower clock than the next
            t0 = time.perf_counter()
            Qd, H4, brk4 = sla.arnoldi(A4, b4, 30)
            barrier()
            cyc4.append(max_over_ranks(time.perf_counter() - t0))
            s4 = min(s4, cyc4[-1])
        if rank == 0:
            extra["arnoldi_cfg4_clocks"] = sampler4a.stop()
        extra["arnoldi_cfg4_seconds_per_cycle"] = cyc4
        b4bytes = 30 * spmv_bytes(n4, n4 * k4) + 8400 * n4    # B_arnoldi cycle, SURVEY.md §8(d)
        extra["arnoldi_cfg4_steps_per_s"] = 30 / s4
        extra["arnoldi_cfg4_ms_per_cycle"] = s4 * 1e3
        extra["arnoldi_cfg4_gbs"] = b4bytes / s4 / 1e9
        extra["cfg4_x_exchange_mode"] = getattr(A4, "dist_p2p_mode", 0)
        del Qd
        # GMRES(30), 10 restarts (BASELINE.json config 4): fixed work, 300 Arnoldi steps with one re-orthogonalisation pass each,
        # 10 restart residuals and 10 solution updates; the stopping test is off so that every cycle runs to its end.
        x04 = sla.SpVector.zeroSV(A4.row_starts[rank + 1] - A4.row_starts[rank])
        sla.gmres(A4, b4, x04, restart=30, nits=30, fixed_work=True)          # warm-up cycle
        sg4, samples4, stream4 = float("inf"), [], []
        sampler4 = ClockSampler(local)
        if rank == 0:
            sampler4.start()
        for _ in range(2):                                  # best of two runs of 10 cycles (both reported, wall clock and stream time)
            barrier()
            ctx.timer_start()
            t0 = time.perf_counter()
            xg4, itg4, resg4 = sla.gmres(A4, b4, x04, restart=30, nits=300, fixed_work=True, info=True)
            barrier()
            samples4.append(max_over_ranks(time.perf_counter() - t0))
            stream4.append(max_over_ranks(ctx.timer_stop()) * 1e-3)
            sg4 = min(sg4, samples4[-1])
        if rank == 0:
            extra["gmres_cfg4_clocks"] = sampler4.stop()
        g4bytes = 10 * (31 * spmv_bytes(n4, n4 * k4) + 17096 * n4)      # per cycle: 31 (#>) + two projection passes per step
        extra["gmres_cfg4_cycles_per_s"] = 10 / sg4
        extra["gmres_cfg4_ms_per_cycle"] = sg4 * 1e2
        extra["gmres_cfg4_gbs"] = g4bytes / sg4 / 1e9
        extra["gmres_cfg4_seconds_per_10_cycles"] = samples4
        extra["gmres_cfg4_stream_seconds_per_10_cycles"] = stream4
        extra["gmres_cfg4_iters"] = itg4
        extra["gmres_cfg4_final_residual"] = resg4
        del A4, xg4
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
                reps = 10 if tag == "k16" else 2
                ms5, _ = timed(lambda: A5.matMat(B5, out=C5), reps, 3 if tag == "k16" else 1)
                b5 = 6 * A5.nnz + 4 * (n + 1) + 4 * n * k5        # B_spmm, SURVEY.md §8(d): bf16 A values, B and C once
                extra[f"spmm_cfg5_{tag}_ms"] = ms5
                extra[f"spmm_cfg5_{tag}_gbs"] = b5 / (ms5 * 1e-3) / 1e9
                extra[f"spmm_cfg5_{tag}_tflops"] = 2 * A5.nnz * k5 / (ms5 * 1e-3) / 1e12
                del A5
            extra["spmm_cfg5_uniform_note"] = ("uniform columns: adjacent rows share no B rows, so the kernel really moves ~82 GB of gathered B rows "
                                               "(11.6x the 7.08 GB algorithmic count, at the L2 bandwidth) — the low fraction is traffic, not idleness")
            del B5, C5
        else:
            # row-partitioned (##), K16 family: every rank holds its row slice of B; the rows of B the block references are
            # gathered with the x-exchange plan applied to k-wide rows, then the tcgen05 tile path runs on the local block
            try:
                k5 = 128
                A5 = gen(sla.GEN_BLOCK16, n, 32, 0x5EED0005)
                s5 = A5.row_starts
                m5 = s5[rank + 1] - s5[rank]
                B5 = sla.DenseMatrix.generate(m5, k5, 0x5EED0055 + rank, sla.BF16)      # this rank's row slice of B
                C5 = sla.DenseMatrix.zeros(m5, k5, sla.BF16)
                ms5, _ = timed(lambda: A5.matMat(B5, out=C5), 10, 3)
                b5 = 6 * n * 32 + 4 * (n + 1) + 4 * n * k5
                extra["spmm_cfg5_k16_dist_ms"] = ms5
                extra["spmm_cfg5_k16_dist_gbs"] = b5 / (ms5 * 1e-3) / 1e9
                del A5, B5, C5
            except Exception as e:
                extra["spmm_cfg5_dist_error"] = str(e)[:200]

    if "cusparse" in want and world == 1:
        # same-box context: cusparseSpMV (CSR, fp64) on the headline matrix; nothing of the product path uses it
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import cusparse_ref

            cz = cusparse_ref.measure("uniform", reps=10)
            extra["cusparse_spmv_cfg2"] = {kk: vv for kk, vv in cz.items() if kk.startswith("cusparse") or kk == "error"}
        except Exception as e:
            extra["cusparse_spmv_cfg2"] = {"error": str(e)[:200]}
    if os.path.exists(os.path.join(ROOT, "profiles", "r02_gather_paths.jsonl")):
        extra["gather_port_microbench"] = "profiles/r02_gather_paths.jsonl: 0.954 random 8-byte gathers per SM-cycle through LDG (= 277 G/s); cfg 2 needs 320 M per (#>)"

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
                     "traffic_source": load_traffic()[3],
                     "traffic_note": "achieved and traffic are per STEP = one (#>) = launches_per_step launches of the kernel",
                     "peak_source": peak_src, "per_gpu": True,
                     "kernel": "spmv_tile_kernel<1024, EPI_NONE> (one launch per column panel, 2 panels at n = 10M)"},
        "e2e": {"value": e2e_gbs, "unit": "GB/s", "h2d_bytes_per_step": 8 * nloc, "d2h_bytes_per_step": 8 * nloc,
                "ms_per_step": e2e_s * 1e3, "api": "sla_spmv_host (pinned host buffers, per-rank slices)"},
        "gpu_launches": int(launches),
        "parity_check": parity,
        "step_ms": step_stats,
        "x_exchange": xch,
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
    ap.add_argument("--extras", default="sptrsv,banded,cfg3,cfg4,cfg5,cusparse",
                    help="comma list of the context runs to include (sptrsv, banded, cfg3, cfg4, cfg5, cusparse)")
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
