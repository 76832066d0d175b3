#!/bin/bash
# 2 GPUs, final code of the round: multi-rank parity assert + weak / strong scaling lines (short), 2-GPU drop-in test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2y_n2_weak.log 2>&1; echo "weak rc=$?"; grep "multi-rank parity" $O/r2y_n2_weak.log | cut -c1-220; tail -n 1 $O/r2y_n2_weak.log | cut -c1-300
timeout 300 $T bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong --no-cpu-baseline > $O/r2y_n2_strong.log 2>&1; echo "strong rc=$?"; tail -n 1 $O/r2y_n2_strong.log | cut -c1-300
timeout 200 python -m pytest tests/test_dropin_gpu.py -x -q -k "two_gpus" > $O/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 $O/r2y_tests.log
