#!/bin/bash
# One short gpurun call, most important output first: GPU parity tests (triangular sweeps last, under their own
# timeout), the bench line, the reference arm, smoke, then ncu captures of the triangular-sweep kernel.
# Every step runs under `timeout -s KILL` so a spinning kernel cannot hold the box.
set -u
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; } > gpurun_out/gpu.txt 2>&1
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline.txt; }

timeout -s KILL 420 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_trisolve.py --durations=8 > gpurun_out/pytest_gpu.log 2>&1
stamp "pytest (all but trisolve) rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout -s KILL 240 python -m pytest tests/test_gpu_trisolve.py -m gpu -q --durations=5 > gpurun_out/pytest_trisolve.log 2>&1
TRI_RC=$?
stamp "pytest trisolve rc=$TRI_RC"; tail -12 gpurun_out/pytest_trisolve.log
BENCH_FLAGS=""; [ "$TRI_RC" != "0" ] && BENCH_FLAGS="--no-sptrsv"
timeout -s KILL 420 python bench.py --steps ${STEPS:-100} --warmup 10 $BENCH_FLAGS > gpurun_out/bench.json 2> gpurun_out/bench.err
stamp "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 120 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
stamp "bench reference rc=$?"; cat gpurun_out/bench_reference.json
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
stamp "smoke rc=$?"; tail -2 gpurun_out/smoke.log
if [ "${PROFILE:-1}" = "1" ]; then
  timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tri_solve -s 2 -c 2 -f -o gpurun_out/prof_trisolve \
     python scripts/bench_sptrsv.py cfg3 > gpurun_out/ncu_trisolve.log 2>&1
  stamp "ncu trisolve rc=$?"
  timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
  stamp "ncu launch list rc=$?"
fi
ls -la gpurun_out
