mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r41_bench_n2.log 2>&1
tail -n 3 gpurun_out/r41_bench_n2.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r41_bench_n1.log 2>&1
tail -n 1 gpurun_out/r41_bench_n1.log
