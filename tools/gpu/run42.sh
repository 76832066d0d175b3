mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -4 > gpurun_out/r42_tests.log
cat gpurun_out/r42_tests.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 20"
timeout 300 $B > gpurun_out/r42_c3.log 2>&1
tail -n 1 gpurun_out/r42_c3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'], d['clocks'])"
