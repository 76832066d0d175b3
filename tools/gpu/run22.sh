mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gibbs_gpu.py tests/test_dropin_gpu.py -m gpu -q -x --timeout 300 -k "gibbs or chains" 2>&1 | tail -5 > gpurun_out/r23_tests.log
timeout 1500 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 60 --nsamples 2 --chains 1 > gpurun_out/r23_gibbs_1.log 2>&1
timeout 1500 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 50 --nsamples 64 --chains 8 > gpurun_out/r23_gibbs_8.log 2>&1
