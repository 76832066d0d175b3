mkdir -p gpurun_out
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x -k test_expected_weights --timeout 900 2>&1 | grep -v "^=========$" | head -80 > gpurun_out/san.log
