#!/bin/bash
# new Gibbs sampler, cooperative K1/K3, -b; C4 / MODEL bench lines; default bench with e2e phases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -x -q > gpurun_out/r2f_tests1.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests1.log; tail -6 gpurun_out/r2f_tests1.log
timeout 600 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_acceptance_gpu.py -x -q > gpurun_out/r2f_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_tests2.log; tail -4 gpurun_out/r2f_tests2.log
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 --gibbs-reads 1000000 > gpurun_out/r2f_C4_1m.log 2>&1; tail -c 1800 gpurun_out/r2f_C4_1m.log
RSEM_B200_TIMING=1 timeout 600 python bench.py --workload C4 > gpurun_out/r2f_C4_10m.log 2>&1; tail -c 2200 gpurun_out/r2f_C4_10m.log
timeout 600 python bench.py --workload MODEL > gpurun_out/r2f_MODEL.log 2>&1; tail -c 1600 gpurun_out/r2f_MODEL.log
RSEM_B200_NO_COOP=1 timeout 600 python bench.py --workload MODEL > gpurun_out/r2f_MODEL_nocoop.log 2>&1; tail -c 700 gpurun_out/r2f_MODEL_nocoop.log
timeout 900 python bench.py > gpurun_out/r2f_bench_default.log 2>&1; tail -c 3000 gpurun_out/r2f_bench_default.log
