mkdir -p gpurun_out
./tools/micro/red_bench 21 > gpurun_out/r27_red_bench.log 2>&1
cat gpurun_out/r27_red_bench.log
