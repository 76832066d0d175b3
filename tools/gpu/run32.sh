mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 10 --variant 4"
for cfg in "1024 2" "512 1" "512 3" "512 4"; do set -- $cfg
RSEM_B200_CTA_THREADS=$1 RSEM_B200_ROW_UNROLL=$2 timeout 300 $B > gpurun_out/r32_c3_$1_$2.log 2>&1
echo "T=$1 U=$2"; tail -n 1 gpurun_out/r32_c3_$1_$2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
RSEM_B200_CTA_THREADS=512 ncu --set full --clock-control none --import-source on -k regex:estep_rows -s 3 -c 1 -o gpurun_out/r32_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline --no-e2e --variant 4 > gpurun_out/r32_ncu.log 2>&1
