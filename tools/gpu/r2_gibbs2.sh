#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 > gpurun_out/r2j_C4_1m.log 2>&1
echo "1M: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_C4_1m.log) $(grep -E 'gibbs chain' gpurun_out/r2j_C4_1m.log | tail -1 | cut -c1-200)"
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 > gpurun_out/r2j_C4_10m.log 2>&1
echo "10M: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_C4_10m.log) $(grep -E 'gibbs chain' gpurun_out/r2j_C4_10m.log | tail -1 | cut -c1-200)"
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --gibbs-chains 1 > gpurun_out/r2j_C4_10m_1chain.log 2>&1
echo "10M 1 chain: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2j_C4_10m_1chain.log)"
tail -c 900 gpurun_out/r2j_C4_10m.log
