#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -x -q -k "gibbs or chains" > gpurun_out/r2h_tests.log 2>&1; tail -5 gpurun_out/r2h_tests.log
RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 > gpurun_out/r2h_C4_1m.log 2>&1; grep -E "bench C4|gibbs chain|Error|error" gpurun_out/r2h_C4_1m.log | tail -8; tail -c 1500 gpurun_out/r2h_C4_1m.log | head -c 400
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 > gpurun_out/r2h_C4_10m.log 2>&1; grep -E "bench C4|gibbs chain|Error|error" gpurun_out/r2h_C4_10m.log | tail -8; tail -c 1500 gpurun_out/r2h_C4_10m.log | head -c 400
