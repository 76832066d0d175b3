mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_em_kernels_gpu.py tests/test_dropin_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -4 > gpurun_out/r33_tests.log
cat gpurun_out/r33_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
for T in 1024 512; do
RSEM_B200_CTA_THREADS=$T timeout 300 $B > gpurun_out/r33_c3_$T.log 2>&1
echo "T=$T"; tail -n 1 gpurun_out/r33_c3_$T.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
