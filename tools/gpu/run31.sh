mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -4 > gpurun_out/r31_tests.log
cat gpurun_out/r31_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 10 --variant 4"
for cfg in "256 8" "512 8" "128 8" "256 4" "256 16"; do set -- $cfg
RSEM_B200_CTA_THREADS=$1 RSEM_B200_GROUP=$2 timeout 300 $B > gpurun_out/r31_c3_$1_$2.log 2>&1
echo "T=$1 G=$2"; tail -n 1 gpurun_out/r31_c3_$1_$2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
