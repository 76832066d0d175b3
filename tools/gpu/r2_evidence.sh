#!/bin/bash
# round 2 evidence run: the bench lines of the other BASELINE workloads (C2, C5, C4, MODEL), the launch list of the default
# bench command, one --set full capture of the class-layout K2 at C3 and one of the Gibbs kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python bench.py --workload C2 --no-e2e > $O/r2m_bench_C2.log 2>&1; echo "C2 rc=$?"; tail -n 1 $O/r2m_bench_C2.log | cut -c1-700
timeout 240 python bench.py --workload C5 --no-e2e > $O/r2m_bench_C5.log 2>&1; echo "C5 rc=$?"; tail -n 1 $O/r2m_bench_C5.log | cut -c1-700
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 > $O/r2m_bench_C4.log 2>&1; echo "C4 rc=$?"; tail -n 1 $O/r2m_bench_C4.log | cut -c1-900
timeout 200 python bench.py --workload MODEL > $O/r2m_bench_MODEL.log 2>&1; echo "MODEL rc=$?"; tail -n 1 $O/r2m_bench_MODEL.log | cut -c1-900
timeout 200 ncu --set full --clock-control none --import-source on -k regex:estep_class -s 3 -c 1 -f -o $O/r2m_k2_class_c3 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-traffic > $O/r2m_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2m_launches_c3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-traffic > $O/r2m_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gibbs_parallel -s 1 -c 1 -f -o $O/r2m_gibbs_1m python bench.py --workload C4 --gibbs-reads 1000000 --steps 4 --no-cpu-baseline > $O/r2m_ncu_gibbs.log 2>&1; echo "ncu gibbs rc=$?"
ls -la $O | tail -20
