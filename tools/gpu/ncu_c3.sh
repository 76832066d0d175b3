mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 > gpurun_out/r39_tests.log
cat gpurun_out/r39_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"estep|theta_update|tile_|gather_u64|max_degree|abs_kernel|max_diff" -c 60 --csv --log-file gpurun_out/r39_launches_c3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r39_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_rows -s 3 -c 1 -o gpurun_out/r39_k2_c3 python bench.py --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r39_ncu_full.log 2>&1
ls -la gpurun_out/r39*
