mkdir -p gpurun_out
export RSEM_B200_TIMING_GAPS=1
B="python bench.py --no-cpu-baseline --no-e2e --steps 20"
timeout 300 $B > gpurun_out/r37_c3.log 2>&1
grep "between" gpurun_out/r37_c3.log; tail -n 1 gpurun_out/r37_c3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"estep|theta_update" -c 40 --csv --log-file gpurun_out/r37_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --scale 0.2 > gpurun_out/r37_ncu.log 2>&1
tail -n 12 gpurun_out/r37_launches.csv | cut -c1-250
