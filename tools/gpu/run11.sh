mkdir -p gpurun_out
python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r11_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r11_bench_c3.log 2>&1
RSEM_B200_GROUP=2 python bench.py --no-cpu-baseline > gpurun_out/r11_bench_c3_g2.log 2>&1
RSEM_B200_GROUP=8 python bench.py --no-cpu-baseline > gpurun_out/r11_bench_c3_g8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_tma -s 3 -c 1 -o gpurun_out/r11_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline > gpurun_out/r11_ncu_full.log 2>&1
