mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -5 > gpurun_out/r28_tests.log
python bench.py --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/r28_c3.log 2>&1
cat gpurun_out/r28_tests.log; tail -n 1 gpurun_out/r28_c3.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'])"
