mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 10 --scale 0.2"
RSEM_B200_GROUP=2 $B > gpurun_out/r29_g2.log 2>&1
RSEM_B200_GROUP=8 $B > gpurun_out/r29_g8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep_tma -s 3 -c 1 -o gpurun_out/r29_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r29_ncu.log 2>&1
for f in gpurun_out/r29_g2.log gpurun_out/r29_g8.log; do tail -n 1 $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"; done
