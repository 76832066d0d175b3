mkdir -p gpurun_out
timeout 1500 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 50 --nsamples 64 --chains 8 > gpurun_out/r15_gibbs_1m.log 2>&1
timeout 1500 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 200 --nsamples 8 --chains 1 > gpurun_out/r15_gibbs_1m_1chain.log 2>&1
timeout 1500 python tools/bench_gibbs.py --N1 4000000 --M 50000 --burnin 50 --nsamples 64 --chains 8 > gpurun_out/r15_gibbs_4m.log 2>&1
