mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
for X in 0 1; do
RSEM_B200_THETA_TEX=$X timeout 300 $B > gpurun_out/r34_c3_tex$X.log 2>&1
echo "TEX=$X"; tail -n 1 gpurun_out/r34_c3_tex$X.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
RSEM_B200_THETA_TEX=1 timeout 600 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -3
