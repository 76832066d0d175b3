#!/bin/bash
# class-layout K2 at C3: build-phase timing, parameter sweep, one --set full capture, launch list, default bench (with e2e)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
RSEM_B200_CLASS_TIMING=1 $B > gpurun_out/r2b_c3_default.log 2>&1
for R in 4 16 32; do RSEM_B200_CLASS_ROWS=$R $B > gpurun_out/r2b_c3_R$R.log 2>&1; done
ncu --set full --clock-control none --import-source on -k regex:estep_class -s 3 -c 1 -o gpurun_out/r2b_k2_class_c3 python bench.py --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2b_launches_c3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b_ncu_launches.log 2>&1
python bench.py > gpurun_out/r2b_bench_default.log 2>&1
for f in gpurun_out/r2b_c3_*.log; do echo "$f: $(grep -o '"k2_ms_per_launch": [0-9.]*' $f) $(grep -o '"ms_per_step": [0-9.]*' $f)"; done
grep "class layout" gpurun_out/r2b_c3_default.log
tail -c 2500 gpurun_out/r2b_bench_default.log
