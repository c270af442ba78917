#!/bin/bash
# One short 1-GPU call: the triangular-sweep kernel variants (parity tests under each, then the timing sweep),
# then a compute-sanitizer memcheck pass over a few small parity tests.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_tri.txt; }
SLA_TRI_MODE=1 timeout -s KILL 120 python -m pytest tests/test_gpu_trisolve.py -m gpu -q -x > gpurun_out/pytest_tri_mode1.log 2>&1
stamp "trisolve tests, persistent kernel rc=$?"; tail -3 gpurun_out/pytest_tri_mode1.log
SLA_TRI_MODE=1 SLA_TRI_LIF=1 SLA_TRI_BACKOFF=200 timeout -s KILL 120 python -m pytest tests/test_gpu_trisolve.py -m gpu -q -x > gpurun_out/pytest_tri_mode1b.log 2>&1
stamp "trisolve tests, persistent kernel + back-off rc=$?"; tail -3 gpurun_out/pytest_tri_mode1b.log
SLA_TRI_MODE=0 SLA_TRI_BACKOFF=200 timeout -s KILL 120 python -m pytest tests/test_gpu_trisolve.py -m gpu -q -x > gpurun_out/pytest_tri_mode0b.log 2>&1
stamp "trisolve tests, back-off rc=$?"; tail -3 gpurun_out/pytest_tri_mode0b.log
timeout -s KILL 150 python scripts/bench_sptrsv.py --sweep cfg3 cfg2 > gpurun_out/sptrsv_sweep.json 2> gpurun_out/sptrsv_sweep.err
stamp "sweep rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/sptrsv_sweep.json"))
for k, v in d.items():
    print(f"{k:48s} {v['solve_ms']:10.3f} ms  {v['us_per_level']:8.3f} us/level  {v['gbs']:8.1f} GB/s  same_bits={v['same_bits']}")
PY
tail -3 gpurun_out/sptrsv_sweep.err
timeout -s KILL 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
   -k "spmv_bit_exact_short_rows or spmv_long_rows or column_panels or vector_ops or krylov_trajectory" > gpurun_out/sanitizer_memcheck.log 2>&1
stamp "compute-sanitizer memcheck rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
