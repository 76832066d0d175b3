#!/bin/bash
# full GPU test suite + the other bench workloads + the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log; tail -25 gpurun_out/r2e_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -2 gpurun_out/r2e_smoke.log
(time python bench.py --impl reference) > gpurun_out/r2e_reference.log 2>&1; tail -c 2500 gpurun_out/r2e_reference.log
for w in C1 C2 C5; do python bench.py --workload $w --no-e2e > gpurun_out/r2e_bench_$w.log 2>&1; tail -c 1800 gpurun_out/r2e_bench_$w.log | head -c 1500; echo; done
python bench.py --workload C4 > gpurun_out/r2e_bench_C4.log 2>&1; tail -c 2500 gpurun_out/r2e_bench_C4.log
python bench.py --workload MODEL > gpurun_out/r2e_bench_MODEL.log 2>&1; tail -c 2500 gpurun_out/r2e_bench_MODEL.log
