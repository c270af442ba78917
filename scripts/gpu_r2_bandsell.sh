#!/bin/bash
# Round-2 shot for the sliced-ELL band plan (1 GPU, ≈ 60 s): parity tests of both band plans, A/B of the launch shapes against the tile
# kernel at full size (bit-identity checked on all 10M rows), one ncu full capture of the fastest shape.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/timeline_bandsell.txt; }
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "band_plan" > gpurun_out/pytest_bandsell.log 2>&1
stamp "pytest rc=$?"; tail -3 gpurun_out/pytest_bandsell.log
timeout -s KILL 240 python scripts/prof_bandsell.py > gpurun_out/bandsell_ab3.jsonl 2> gpurun_out/bandsell_ab3.err
stamp "A/B rc=$?"; cat gpurun_out/bandsell_ab3.jsonl; tail -3 gpurun_out/bandsell_ab3.err
BEST=$(python - <<'PY'
import json
best, bms = 3, 1e9
for l in open("gpurun_out/bandsell_ab3.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    if d.get("variant") and d["ms"] < bms and d.get("bit_identical_to_tile_kernel"): best, bms = d["variant"], d["ms"]
print(best)
PY
)
SLA_BSELL_VARIANT=$BEST timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:spmv_bsell_kernel -s 30 -c 1 -f -o gpurun_out/prof_spmv_bsell3 \
   python scripts/prof_bandsell.py quick > gpurun_out/ncu_bsell3.log 2>&1
stamp "ncu (variant $BEST) rc=$?"; tail -2 gpurun_out/ncu_bsell3.log
