#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): parity check under every transport, then the bench line with the automatic
# transports and with NCCL for comparison.  Every step under timeout -s KILL.
set -u
N=${N:-2}
TAG=${TAG:-r2}
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_dist_${TAG}_n$N.txt; }
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29520
for variant in ${CHECKS:-auto x0 nccl}; do
  case $variant in
    none) continue;;
    auto) ENVS="";;
    x0)   ENVS="SLA_P2P_X=0";;
    x1)   ENVS="SLA_P2P_X=1";;
    x4)   ENVS="SLA_P2P_X=4";;
    x5)   ENVS="SLA_P2P_X=5 SLA_P2P_ARRIVAL_ALWAYS=1";;      # ALWAYS: the small check matrices take the phased path too
    x5p*c*) v=${variant#x5p}; ENVS="SLA_P2P_X=5 SLA_P2P_PANELS=${v%%c*} SLA_P2P_PUSH_CTAS=${v##*c}";;   # e.g. x5p1,1,2c64
    arrival) ENVS="SLA_P2P_X=2 SLA_P2P_ARRIVAL_ALWAYS=1";;
    noinline) ENVS="SLA_P2P_INLINE=0";;
    nccl) ENVS="SLA_P2P=0";;
  esac
  port=$((port+1))
  env $ENVS timeout -s KILL 240 $RUN --master-port $port tests/dist_check.py > gpurun_out/dist_check_${TAG}_${variant}_n$N.log 2>&1
  stamp "dist_check $variant rc=$?"; grep -A12 DIST_CHECK gpurun_out/dist_check_${TAG}_${variant}_n$N.log | head -16; tail -3 gpurun_out/dist_check_${TAG}_${variant}_n$N.log | cut -c1-300
done
for variant in ${BENCHES:-auto x0}; do
  case $variant in
    none) continue;;
    auto) ENVS="";;
    x0)   ENVS="SLA_P2P_X=0";;
    x4)   ENVS="SLA_P2P_X=4";;
    x5)   ENVS="SLA_P2P_X=5";;
    x5p*c*) v=${variant#x5p}; ENVS="SLA_P2P_X=5 SLA_P2P_PANELS=${v%%c*} SLA_P2P_PUSH_CTAS=${v##*c}";;   # e.g. x5p1,1,2c64
    arrival) ENVS="SLA_P2P_X=2 SLA_P2P_ARRIVAL_ALWAYS=1";;
    nccl) ENVS="SLA_P2P=0";;
  esac
  port=$((port+1))
  env $ENVS timeout -s KILL 400 $RUN --master-port $port bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 --extras ${EXTRAS:-cfg3,cfg4,cfg5} > gpurun_out/bench_${TAG}_${variant}_n$N.json 2> gpurun_out/bench_${TAG}_${variant}_n$N.err
  stamp "bench $variant rc=$?"; cat gpurun_out/bench_${TAG}_${variant}_n$N.json; tail -4 gpurun_out/bench_${TAG}_${variant}_n$N.err | cut -c1-300
done
