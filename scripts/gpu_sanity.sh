#!/bin/bash
# Smallest possible 1-GPU check of the built library: smoke + the two parity files, fail-fast.
set -u
mkdir -p gpurun_out
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanity_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/sanity_smoke.log
timeout -s KILL 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trisolve.py tests/test_matrix_market.py -m gpu -q -x > gpurun_out/sanity_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/sanity_pytest.log
