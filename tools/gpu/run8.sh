mkdir -p gpurun_out
python bench.py > gpurun_out/r8_bench_c3.log 2>&1
python bench.py --impl reference > gpurun_out/r8_bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"estep|theta_update|tile_bounds|gather_u64|max_degree|abs_kernel|max_diff" --csv --log-file gpurun_out/r8_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_tma -s 5 -c 1 -o gpurun_out/r8_k2_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r8_ncu_full.log 2>&1
python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r8_bench_c2.log 2>&1
python bench.py --workload C1 --no-cpu-baseline > gpurun_out/r8_bench_c1.log 2>&1
