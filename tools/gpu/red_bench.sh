mkdir -p gpurun_out
./tools/micro/red_bench 21 > gpurun_out/red_bench.log 2>&1
cat gpurun_out/red_bench.log
