mkdir -p gpurun_out
python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r9_tests.log
python bench.py --no-cpu-baseline --variant 3 > gpurun_out/r9_bench_c3_warp.log 2>&1
python bench.py --no-cpu-baseline --variant 1 > gpurun_out/r9_bench_c3_cta.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_warp -s 3 -c 1 -o gpurun_out/r9_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline --variant 3 > gpurun_out/r9_ncu_full.log 2>&1
