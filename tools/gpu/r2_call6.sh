#!/bin/bash
# Gibbs: lean row path on / off on the same box, device time of the kernels next to the wall clock of gibbs_run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for fast in 1 0 1; do
  RSEM_B200_GIBBS_FAST=$fast RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 --no-cpu-baseline > $O/r2t_C4_1m_fast$fast.log 2>&1
  echo "fast=$fast 1M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2t_C4_1m_fast$fast.log)"; grep -E "gibbs kernels|gibbs_run" $O/r2t_C4_1m_fast$fast.log | tail -4 | cut -c1-160
done
for fast in 1 0; do
  RSEM_B200_GIBBS_FAST=$fast RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --no-cpu-baseline > $O/r2t_C4_10m_fast$fast.log 2>&1
  echo "fast=$fast 10M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2t_C4_10m_fast$fast.log)"; grep -E "gibbs kernels|gibbs_run" $O/r2t_C4_10m_fast$fast.log | tail -4 | cut -c1-160
done
