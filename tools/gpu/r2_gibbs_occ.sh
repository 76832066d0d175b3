#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for o in 2 3 4; do
  RSEM_B200_GIBBS_OCC=$o RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 --no-cpu-baseline > gpurun_out/r2i_C4_1m_occ$o.log 2>&1
  echo "occ $o 1M: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2i_C4_1m_occ$o.log) $(grep -E 'gibbs chain' gpurun_out/r2i_C4_1m_occ$o.log | tail -1 | cut -c1-60)"
  RSEM_B200_GIBBS_OCC=$o RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --no-cpu-baseline > gpurun_out/r2i_C4_10m_occ$o.log 2>&1
  echo "occ $o 10M: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2i_C4_10m_occ$o.log)"
done
RSEM_B200_GIBBS_OCC=4 timeout 300 python -m pytest tests/test_gibbs_gpu.py -x -q 2>&1 | tail -2
