mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 > gpurun_out/r46_tests.log
cat gpurun_out/r46_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r46_smoke.log 2>&1; tail -n 2 gpurun_out/r46_smoke.log
( time python bench.py ) > gpurun_out/r46_bench_c3.log 2>&1
tail -n 4 gpurun_out/r46_bench_c3.log | cut -c1-1500
