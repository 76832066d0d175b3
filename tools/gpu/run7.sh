mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15 > gpurun_out/r7_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r7_bench_c3.log 2>&1
