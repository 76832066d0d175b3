mkdir -p gpurun_out
python -m pytest tests/test_dropin_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -8 > gpurun_out/r13_tests_2gpu.log
cat gpurun_out/r13_tests_2gpu.log
