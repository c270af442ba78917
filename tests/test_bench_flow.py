"""CPU test of bench.py's control flow and JSON contract with a FAKE backend (no kernels run, numbers are dummies):
every extras block is walked, the line parses, and the keys the driver reads are present.  The real numbers come from
the GPU box; this only guards the harness against NameErrors / contract drift."""
import io
import json
import sys
import types
from contextlib import redirect_stdout

import numpy as np


class _Vec:
    def __init__(self, n):
        self.dim = n

    def toDenseListSV(self):
        return np.zeros(self.dim)

    def copy(self):
        return _Vec(self.dim)

    def norm2(self):
        return 1.0

    def __sub__(self, other):
        return _Vec(self.dim)


class _Mat:
    nnz = 320_000_000
    spmv_bytes = 1

    def __init__(self, n):
        self.n, self.h, self.row_starts = n, None, [0, n]

    def matVec(self, x, out=None):
        return out or _Vec(self.n)

    def __matmul__(self, x):
        return _Vec(self.n)

    def matMat(self, b, out=None):
        return out

    def triAnalysis(self, upper):
        return 7, 1000


class _Ctx:
    launches = 0
    p2p = False

    class lib:
        @staticmethod
        def sla_spmv_host(*a):
            return 0

    h = None

    def __init__(self, *a, **k):
        self._t = 0

    def sync(self):
        pass

    def check(self, st):
        return st

    def timer_start(self):
        pass

    def timer_stop(self):
        _Ctx.launches += 2
        return 1.5

    def pinned(self, n):
        return np.zeros(n)

    def set_option(self, name, value):
        pass


def _fake_backend():
    sla = types.ModuleType("sparse_linear_algebra_b200")
    sla.Context = _Ctx
    sla.set_default_context = lambda c: None
    for k, name in enumerate(("GEN_UNIFORM", "GEN_BANDED", "GEN_LAPLACE2D", "GEN_BLOCK16")):
        setattr(sla, name, k)
    sla.BF16 = 1
    sla.SpMatrix = types.SimpleNamespace(generate=lambda kind, n, k, seed, band=0: _Mat(n))
    sla.SpVector = types.SimpleNamespace(zeroSV=lambda n: _Vec(n))
    sla.DenseMatrix = types.SimpleNamespace(generate=lambda *a: object(), zeros=lambda *a: object())
    sla.bicgsInit = lambda A, b, x0: types.SimpleNamespace(r=_Vec(b.dim), x=_Vec(b.dim))
    sla.gmres = lambda A, b, x0, **kw: (x0, 300, 1e-9) if kw.get("info") else x0
    sla.bicgstabStep = lambda A, r, st: st
    sla.arnoldi = lambda A, b, kn: (object(), None, False)
    sla.triLowerSolve = lambda M, rhs, out=None: out
    sd = types.ModuleType("sparse_linear_algebra_b200.dist")
    sd.generate_vector_slice = lambda ctx, n, seed, starts, rank: _Vec(starts[rank + 1] - starts[rank])
    sla.dist = sd
    return sla, sd


def test_bench_line_contract(monkeypatch):
    sla, sd = _fake_backend()
    monkeypatch.setitem(sys.modules, "sparse_linear_algebra_b200", sla)
    monkeypatch.setitem(sys.modules, "sparse_linear_algebra_b200.dist", sd)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    import bench

    monkeypatch.setattr(bench, "cpu_baseline", lambda threads, **k: (20.0, 0.03, "sample", 0.6))
    monkeypatch.setattr(bench, "parity_cfg2", lambda *a, **k: {"rows": 300, "bit_exact_rows": 300, "max_err_over_bound": 0.0, "ok": True})
    monkeypatch.setattr(bench, "load_traffic", lambda: (4.2e9, 2.1e9, 2, "fake capture"))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1})
    args = types.SimpleNamespace(gpus=1, steps=5, warmup=3, impl="b200", quick=False, no_cpu=False, no_sptrsv=False,
                                 extras="sptrsv,banded,cfg3,cfg4,cfg5")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_gpu(args)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline", "parity_check", "step_ms"):
        assert key in d, key
    assert d["metric"] == "csr_spmv_fp64_gbs" and d["unit"] == "GB/s" and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert d["roofline"]["traffic"] == d["roofline"]["traffic_per_launch"] * d["roofline"]["launches_per_step"]
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] == 80_000_000
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert d["parity_check"]["ok"] is True and d["extra"]["bicgstab_cfg3_parity"]["ok"] is True
    for key in ("spmv_banded_gbs", "bicgstab_cfg3_iters_per_s", "arnoldi_cfg4_steps_per_s", "gmres_cfg4_cycles_per_s", "spmm_cfg5_k16_ms", "sptrsv_cfg3_lower_ms"):
        assert key in d["extra"], key


def test_reference_arm_contract(monkeypatch, ora):
    import bench

    monkeypatch.delenv("RANK", raising=False)
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=3, impl="reference")
    # a small sample keeps the CPU test short; the shipped default is config 2 itself (10 M rows) when the host memory allows
    real_synth = ora.SpMatrix.synth
    monkeypatch.setattr(ora.SpMatrix, "synth", staticmethod(lambda kind, n, k, seed, band=0: real_synth(kind, 20000, k, seed, band)))
    real_vsynth = ora.SpVector.synth
    monkeypatch.setattr(ora.SpVector, "synth", staticmethod(lambda seed, n: real_vsynth(seed, 20000)))
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    d = json.loads(buf.getvalue().strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "csr_spmv_fp64_gbs" and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
