#!/bin/bash
# Multi-GPU call (gpurun --gpus N): parity check and a short bench with the peer-memory collectives (default) and
# with NCCL (SLA_P2P=0).  P2PX=2 selects the arrival-order exchange; EXPERIMENTAL=1 (default here) also checks the row-partitioned (##).  SLA_P2P_X=1 also routes the x exchange through peer memory (opt-in).  Every step under timeout -s KILL.
set -u
N=${N:-2}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_p2p.txt; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SLA_DIST_CHECK_EXPERIMENTAL=${EXPERIMENTAL:-1} SLA_P2P_X=${P2PX:-1} timeout -s KILL 200 $RUN --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_p2p_n$N.log 2>&1
stamp "dist_check p2p rc=$?"; grep -A12 DIST_CHECK gpurun_out/dist_check_p2p_n$N.log | head -20; tail -5 gpurun_out/dist_check_p2p_n$N.log
SLA_P2P_X=${P2PX:-1} timeout -s KILL 200 $RUN --master-port 29512 bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --extras ${EXTRAS:-cfg3} > gpurun_out/bench_p2p_n$N.json 2> gpurun_out/bench_p2p_n$N.err
stamp "bench p2p rc=$?"; cat gpurun_out/bench_p2p_n$N.json; tail -5 gpurun_out/bench_p2p_n$N.err
SLA_P2P=0 timeout -s KILL 200 $RUN --master-port 29513 bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --extras ${EXTRAS:-cfg3} > gpurun_out/bench_nccl_n$N.json 2> gpurun_out/bench_nccl_n$N.err
stamp "bench nccl rc=$?"; cat gpurun_out/bench_nccl_n$N.json; tail -3 gpurun_out/bench_nccl_n$N.err
