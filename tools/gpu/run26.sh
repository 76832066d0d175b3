mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
$B > gpurun_out/r26_base.log 2>&1
RSEM_B200_TILE_ORDER=1 $B > gpurun_out/r26_contig.log 2>&1
$B --sort-rows > gpurun_out/r26_sorted.log 2>&1
RSEM_B200_TILE_ORDER=1 $B --sort-rows > gpurun_out/r26_sorted_contig.log 2>&1
tail -n 2 gpurun_out/r26_*.log
