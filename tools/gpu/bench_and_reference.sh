mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r40_bench_c3.log 2>&1
( time python bench.py --impl reference ) > gpurun_out/r40_bench_ref.log 2>&1
tail -n 5 gpurun_out/r40_bench_c3.log; tail -n 5 gpurun_out/r40_bench_ref.log
