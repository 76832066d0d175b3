mkdir -p gpurun_out
python tools/bench_dropin.py --read-type 1 --N1 1000000 --M 50000 --avg-family 10 --read-len 100 --rounds 20 > gpurun_out/r12_dropin_seq.log 2>&1
python tools/bench_dropin.py --read-type 3 --N1 500000 --M 50000 --avg-family 20 --read-len 100 --rounds 20 --est-rspd 1 > gpurun_out/r12_dropin_peq.log 2>&1
python tools/bench_dropin.py --read-type 0 --N1 100000 --M 5000 --avg-family 5 --read-len 50 --rounds 20 > gpurun_out/r12_dropin_c1.log 2>&1
