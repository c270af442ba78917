#!/bin/bash
export NP=${NP:-4}
run() { echo -n "$1: "; env $2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $NP --steps 40 --warmup 5 --quick 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'GB/s')"; }
run "allgather (no pipeline)" "SLA_DIST_NO_PIPELINE=1" 29551
run "pipelined P=2" "X=1" 29552
run "pipelined P=4" "SLA_DIST_PANELS=4" 29553
