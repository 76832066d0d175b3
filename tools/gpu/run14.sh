mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -m gpu -q -x --timeout 300 -k "gibbs or chains" 2>&1 | tail -25 > gpurun_out/r14_tests.log
cat gpurun_out/r14_tests.log
