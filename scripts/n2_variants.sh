#!/bin/bash
# 2-GPU SpMV exchange variants (cfg 2, strong scaling): ms per (#>)
run() { echo -n "$1: "; env $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port $3 bench.py --gpus ${NP:-2} --steps 40 --warmup 5 --quick 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'GB/s')"; }
run "allgather (no pipeline)" "SLA_DIST_NO_PIPELINE=1" 29541
run "pipelined P=2" "X=1" 29542
run "pipelined P=2, p2p channels 32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" 29543
run "allgather, NCCL_MIN_NCHANNELS=32" "SLA_DIST_NO_PIPELINE=1 NCCL_MIN_NCHANNELS=32" 29544
