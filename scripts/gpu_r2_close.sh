#!/bin/bash
# Round-2 closing check (1 GPU, ≈ 70 s): the bench line as the driver runs it (now with the sliced-ELL band measurement), then the GPU tests.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_close.txt; }
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_close.json 2> gpurun_out/bench_close.err
stamp "bench rc=$?"; cut -c1-300 gpurun_out/bench_close.json; tail -3 gpurun_out/bench_close.err
timeout -s KILL 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_close.log 2>&1
stamp "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_close.log
