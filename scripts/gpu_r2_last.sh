#!/bin/bash
# Round-2 closing call (1 GPU, ≈ 100 s): parity tests (with the rotated-panel hook test), smoke, the bench line as the driver runs it,
# and the one-GPU push-contention probe behind DESIGN.md section 5.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_last.txt; }
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_last.log 2>&1
stamp "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_last.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_last.log 2>&1
stamp "smoke rc=$?"; tail -1 gpurun_out/smoke_last.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
stamp "bench rc=$?"; cut -c1-400 gpurun_out/bench_last.json; tail -3 gpurun_out/bench_last.err
timeout -s KILL 200 python scripts/prof_push_contention.py > gpurun_out/push_contention.jsonl 2> gpurun_out/push_contention.err
stamp "push contention rc=$?"; cat gpurun_out/push_contention.jsonl | cut -c1-1500; tail -3 gpurun_out/push_contention.err
