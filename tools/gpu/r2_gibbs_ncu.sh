#!/bin/bash
# --set full capture (with source correlation) of the final Gibbs kernel, C4 at 1 M reads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gibbs_parallel -s 1 -c 1 -f -o gpurun_out/r2x_gibbs_1m python bench.py --workload C4 --gibbs-reads 1000000 --steps 4 --no-cpu-baseline > gpurun_out/r2x_ncu_gibbs.log 2>&1; echo "ncu gibbs rc=$?"
