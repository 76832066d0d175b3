#!/bin/bash
# Gibbs fast path for register-resident rows: parity tests + C4; e2e phases of every job
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -x -q -k "gibbs or chains or family" > $O/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 $O/r2s_tests.log
RSEM_B200_TIMING=1 timeout 200 python bench.py --workload C4 --gibbs-reads 1000000 --no-cpu-baseline > $O/r2s_C4_1m.log 2>&1
echo "1M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2s_C4_1m.log) $(grep -E 'gibbs chain' $O/r2s_C4_1m.log | tail -1 | cut -c1-220)"
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 > $O/r2s_C4_10m.log 2>&1
echo "10M: $(grep -o '"ms_per_step": [0-9.]*' $O/r2s_C4_10m.log) $(grep -E 'gibbs chain' $O/r2s_C4_10m.log | tail -1 | cut -c1-220)"
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --gibbs-chains 1 --no-cpu-baseline > $O/r2s_C4_10m_1chain.log 2>&1
echo "10M 1 chain: $(grep -o '"ms_per_step": [0-9.]*' $O/r2s_C4_10m_1chain.log) $(grep -E 'gibbs chain' $O/r2s_C4_10m_1chain.log | tail -1 | cut -c1-120)"
timeout 300 python bench.py --no-cpu-baseline --no-traffic > $O/r2s_bench_e2e.log 2>&1; echo "bench rc=$?"
tail -n 1 $O/r2s_bench_e2e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e'])[:1400])"
