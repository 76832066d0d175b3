#!/bin/bash
# whole-executable wall clock at C1 (100 k reads x 5 k transcripts, 20 rounds) against the reference with 8 threads
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
RSEM_B200_TIMING=1 timeout 100 python tools/bench_dropin.py --read-type 0 --N1 100000 --M 5000 --avg-family 5 --read-len 50 --rounds 20 --ref-threads 8 > gpurun_out/r2v_c1_exe.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v_c1_exe.log | cut -c1-900
