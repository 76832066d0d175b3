mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/r2_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c3.log 2>&1
RSEM_B200_GROUP=8 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c3_g8.log 2>&1
RSEM_B200_GROUP=32 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c3_g32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_tma -s 3 -c 1 -o gpurun_out/r2_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
