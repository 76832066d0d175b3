#!/bin/bash
# 2 GPUs: NVLink peer-memory count reduction (p2p.cu), multi-rank parity assert, weak + strong scaling, 2-GPU drop-in tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2k_n2_weak.log 2>&1; echo "weak rc=$?"; tail -n 1 gpurun_out/r2k_n2_weak.log | cut -c1-1500
timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong > gpurun_out/r2k_n2_strong.log 2>&1; echo "strong rc=$?"; tail -n 1 gpurun_out/r2k_n2_strong.log | cut -c1-1500
RSEM_B200_NO_P2P=1 timeout 400 $T bench.py --gpus 2 --steps 20 --warmup 3 --scaling strong --no-e2e > gpurun_out/r2k_n2_strong_nccl.log 2>&1; echo "strong nccl rc=$?"; tail -n 1 gpurun_out/r2k_n2_strong_nccl.log | cut -c1-600
timeout 300 $T bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --ref-budget 40 > gpurun_out/r2k_n2_ref.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/r2k_n2_ref.log | cut -c1-400
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or 2gpu or gibbs" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r2k_tests.log
