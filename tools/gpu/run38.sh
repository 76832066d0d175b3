mkdir -p gpurun_out
export RSEM_B200_TILE_ORDER=0
B="python bench.py --no-cpu-baseline --no-e2e --steps 10 --scale 0.4"
timeout 300 $B > gpurun_out/r38_base.log 2>&1
timeout 300 $B --sort-rows deg > gpurun_out/r38_deg.log 2>&1
for f in base deg; do echo $f; tail -n 1 gpurun_out/r38_$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"; done
