# compute-sanitizer memcheck over one kernel test and one paired-end drop-in run (K1 / K2 / K3 / K4 incl. the byte-stream
# loaders that read up to 15 bytes into the arrays' tail padding)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_em_kernels_gpu.py -m gpu -q -x -k "test_expected_weights" --timeout 900 2>&1 | grep -v "^=========$" | tail -12 > gpurun_out/san_kernels.log
python - <<'PY'
import sys
sys.path.insert(0, 'tests')
import rsem_files as rf
rf.gen_dataset('/tmp/gs', read_type=3, M=200, N1=3000, N0=100, avg_family=5, read_len=50, seed=3)
rf.gen_dataset('/tmp/gs1', read_type=1, M=200, N1=3000, N0=100, avg_family=5, read_len=50, seed=4, est_rspd=1)
PY
export RSEM_MAX_ROUND=13 RSEM_MIN_ROUND=13
( cd /tmp/gs && timeout 600 compute-sanitizer --tool memcheck --print-limit 5 $GRAFT_REPO_ROOT/bin/rsem-run-em ref/r 3 s s.temp/s s.stat/s -p 4 --gibbs-out -q 2>&1 | tail -6 ) > gpurun_out/san_em_pe.log
( cd /tmp/gs1 && timeout 600 compute-sanitizer --tool memcheck --print-limit 5 $GRAFT_REPO_ROOT/bin/rsem-run-em ref/r 1 s s.temp/s s.stat/s -p 4 --gibbs-out -q 2>&1 | tail -6 ) > gpurun_out/san_em_se.log
( cd /tmp/gs1 && timeout 600 compute-sanitizer --tool memcheck --print-limit 5 $GRAFT_REPO_ROOT/bin/rsem-run-gibbs ref/r s.temp/s s.stat/s 20 10 1 -p 2 --seed 5 -q 2>&1 | tail -6 ) > gpurun_out/san_gibbs.log
tail -n 4 gpurun_out/san_*.log
