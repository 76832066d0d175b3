#!/bin/bash
# Gibbs component walk variants (RSEM_B200_GIBBS_PF = 2 converged + pipelined, 1 pipelined, 0 row-at-a-time): parity tests,
# C4 at 1 M and 10 M reads, one --set full capture with source correlation
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -x -q -k "gibbs or chains or family" > $O/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 $O/r2p_tests.log
for pf in 2 1; do
  RSEM_B200_GIBBS_PF=$pf RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 --no-cpu-baseline > $O/r2p_C4_1m_pf$pf.log 2>&1
  echo "pf=$pf 1M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2p_C4_1m_pf$pf.log) $(grep -E 'gibbs chain' $O/r2p_C4_1m_pf$pf.log | tail -1 | cut -c1-220)"
  RSEM_B200_GIBBS_PF=$pf RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --no-cpu-baseline > $O/r2p_C4_10m_pf$pf.log 2>&1
  echo "pf=$pf 10M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2p_C4_10m_pf$pf.log) $(grep -E 'gibbs chain' $O/r2p_C4_10m_pf$pf.log | tail -1 | cut -c1-220)"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gibbs_parallel -s 1 -c 1 -f -o $O/r2p_gibbs_1m python bench.py --workload C4 --gibbs-reads 1000000 --steps 4 --no-cpu-baseline > $O/r2p_ncu_gibbs.log 2>&1; echo "ncu gibbs rc=$?"
