#!/bin/bash
# Round-2 single-GPU validation: all GPU parity tests, the bench line, the reference arm, smoke.  TAG names the outputs.
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_$TAG.txt; }
timeout -s KILL 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
stamp "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
if [ "${BENCH:-1}" = "1" ]; then
  timeout -s KILL 600 python bench.py --steps ${STEPS:-100} --warmup 10 ${BENCH_FLAGS:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  stamp "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
fi
if [ "${REFERENCE:-0}" = "1" ]; then
  timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_$TAG.json 2>> gpurun_out/bench_$TAG.err
  stamp "bench reference rc=$?"; cat gpurun_out/bench_reference_$TAG.json
fi
