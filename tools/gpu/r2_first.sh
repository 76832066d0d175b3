#!/bin/bash
# round 2, first GPU call: class-layout K2 correctness + first timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests/test_em_kernels_gpu.py -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
compute-sanitizer --tool memcheck python -m pytest tests/test_em_kernels_gpu.py -x -q -k "class_layout_matches_oracle and (case0 or case3 or case6 or case7)" > gpurun_out/r2a_sanitizer.log 2>&1
tail -3 gpurun_out/r2a_sanitizer.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_c3_class1024.log 2>&1
RSEM_B200_CLASS_THREADS=512 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_c3_class512.log 2>&1
RSEM_B200_NO_CLASS=1 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_c3_rows.log 2>&1
tail -c 1500 gpurun_out/r2a_c3_class1024.log; echo; tail -c 600 gpurun_out/r2a_c3_class512.log; echo; tail -c 600 gpurun_out/r2a_c3_rows.log
