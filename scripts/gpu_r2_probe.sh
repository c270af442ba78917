#!/bin/bash
# Round-2 call 1 (1 GPU): baseline parity run, the gather-port micro-benchmark, cuSPARSE on the bench matrices, banded L1 experiment.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_r2_probe.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for mb in 32 40 80; do timeout -s KILL 60 scripts/microbench/_build/gather_paths $mb 256 >> gpurun_out/gather_paths.jsonl 2>> gpurun_out/gather_paths.err; done
stamp "gather_paths rc=$?"; cat gpurun_out/gather_paths.jsonl
timeout -s KILL 300 python scripts/cusparse_ref.py uniform banded cfg4 cfg3 > gpurun_out/cusparse_ref.json 2> gpurun_out/cusparse_ref.err
stamp "cusparse rc=$?"; cat gpurun_out/cusparse_ref.json; tail -3 gpurun_out/cusparse_ref.err
for h in 3 7; do SLA_SPMV_HINTS=$h timeout -s KILL 100 python scripts/prof_case.py banded 10000000 32 65536 2>&1 | tail -1 | sed "s/^/HINTS=$h /" | tee -a gpurun_out/banded_hints.txt; done
for h in 3 7; do SLA_SPMV_HINTS=$h timeout -s KILL 100 python scripts/prof_case.py uniform 10000000 32 0 2>&1 | tail -1 | sed "s/^/HINTS=$h /" | tee -a gpurun_out/banded_hints.txt; done
stamp "hints done"
timeout -s KILL 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
stamp "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
