mkdir -p gpurun_out
for T in 128 256 512; do
RSEM_B200_CTA_THREADS=$T timeout 600 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -2 > gpurun_out/r30_tests_$T.log
RSEM_B200_CTA_THREADS=$T python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/r30_c3_$T.log 2>&1
cat gpurun_out/r30_tests_$T.log; tail -n 1 gpurun_out/r30_c3_$T.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
