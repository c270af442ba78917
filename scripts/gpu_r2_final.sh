#!/bin/bash
# Round-2 single-GPU record: parity tests, bench line, reference arm, smoke, ncu captures (cfg-2 (#>) for roofline.traffic, the
# pipelined tcgen05 (##) kernel, the Arnoldi kernels) and the launch list of a short bench.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_final.txt; }
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1
stamp "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_final.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
stamp "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout -s KILL 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
stamp "bench rc=$?"; cut -c1-600 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout -s KILL 400 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference_final.json 2>> gpurun_out/bench_final.err
stamp "bench reference rc=$?"; cat gpurun_out/bench_reference_final.json
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmv_tile_kernel -s 6 -c 2 -f -o gpurun_out/prof_spmv_cfg2 \
   python scripts/prof_case.py uniform 10000000 32 0 4 > gpurun_out/ncu_spmv.log 2>&1
stamp "ncu spmv rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmm_bsr_tc_pipe -s 2 -c 1 -f -o gpurun_out/prof_spmm_pipe \
   python scripts/bench_spmm.py block16 10000000 > gpurun_out/ncu_spmm.log 2>&1
stamp "ncu spmm rc=$?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"tsmv_t_kernel|lincomb_kernel" -s 36 -c 2 -f -o gpurun_out/prof_arnoldi2 \
   python scripts/prof_arnoldi.py 4000000 64 30 1 > gpurun_out/ncu_arnoldi2.log 2>&1
stamp "ncu arnoldi rc=$?"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --extras cfg3,cfg4 > gpurun_out/ncu_launches_final.log 2>&1
stamp "ncu launch list rc=$?"
ls -la gpurun_out | tail -15
