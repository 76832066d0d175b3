mkdir -p gpurun_out
timeout 1500 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 60 --nsamples 2 --chains 1 > gpurun_out/r21_gibbs.log 2>&1
