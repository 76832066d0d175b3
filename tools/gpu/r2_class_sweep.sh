#!/bin/bash
# class-layout K2: kernel tests + CTA-shape sweep at C3 + one --set full capture + the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests/test_em_kernels_gpu.py -x -q > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-traffic"
for T in 512 513 544 640 768 1024; do RSEM_B200_CLASS_THREADS=$T $B > gpurun_out/r2c_c3_T$T.log 2>&1; done
for f in gpurun_out/r2c_c3_*.log; do echo "$f: $(grep -o '"k2_ms_per_launch": [0-9.]*' $f) $(grep -o '"ms_per_step": [0-9.]*' $f)"; done
ncu --set full --clock-control none --import-source on -k regex:estep_class -s 3 -c 1 -o gpurun_out/r2c_k2_class_c3 python bench.py --steps 3 --no-cpu-baseline --no-e2e --no-traffic > gpurun_out/r2c_ncu_full.log 2>&1
python bench.py > gpurun_out/r2c_bench_default.log 2>&1
tail -c 3000 gpurun_out/r2c_bench_default.log
python -m pytest tests/test_baseline_sizes_gpu.py -x -q --durations=8 > gpurun_out/r2c_baseline_tests.log 2>&1; tail -15 gpurun_out/r2c_baseline_tests.log
