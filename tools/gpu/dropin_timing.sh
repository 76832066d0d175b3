mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/r25_tests.log
python - <<'PY'
import sys
sys.path.insert(0, 'tests')
import rsem_files as rf
rf.gen_dataset('/tmp/gb', read_type=1, M=50000, N1=2000000, N0=100000, avg_family=10, read_len=100, seed=11)
PY
cd /tmp/gb
export RSEM_MAX_ROUND=20 RSEM_MIN_ROUND=20 RSEM_B200_TIMING=1
( time $GRAFT_REPO_ROOT/bin/rsem-run-em ref/r 1 s s.temp/s s.stat/s -p 32 --gibbs-out -q ) > $GRAFT_REPO_ROOT/gpurun_out/r25_em_timing.log 2>&1
( time $GRAFT_REPO_ROOT/bin/rsem-run-gibbs ref/r s.temp/s s.stat/s 50 64 1 -p 8 --seed 3 -q ) >> $GRAFT_REPO_ROOT/gpurun_out/r25_em_timing.log 2>&1
cp s.stat/s.theta /tmp/ours.theta; cp s.temp/s.ofg /tmp/ours.ofg; cp s.stat/s.model /tmp/ours.model
$GRAFT_REPO_ROOT/oracle/_ref/rsem-build-read-index 32 1 1 s.temp/s_alignable.fq
( time $GRAFT_REPO_ROOT/oracle/_ref/rsem-run-em-rounds ref/r 1 s2 s.temp/s s.stat/s -p 128 --gibbs-out -q ) >> $GRAFT_REPO_ROOT/gpurun_out/r25_em_timing.log 2>&1
python - >> $GRAFT_REPO_ROOT/gpurun_out/r25_em_timing.log 2>&1 <<'PY'
import numpy as np
a = np.array(open('/tmp/ours.theta').read().split()[1:], dtype=float)
b = np.array(open('/tmp/gb/s.stat/s.theta').read().split()[1:], dtype=float)
print('theta max rel err', np.max(np.abs(a-b)/np.maximum(b,1e-30)))
x = open('/tmp/ours.ofg').read().split(); y = open('/tmp/gb/s.temp/s.ofg').read().split()
print('ofg tokens', len(x), len(y), 'identical', x == y)
if x != y:
    xa = np.array(x[:200000], dtype=float); ya = np.array(y[:200000], dtype=float)
    print('ofg first 200k max rel', np.max(np.abs(xa-ya)/np.maximum(np.abs(ya),1e-300)))
print('model identical', open('/tmp/ours.model').read() == open('/tmp/gb/s.stat/s.model').read())
PY
cat $GRAFT_REPO_ROOT/gpurun_out/r25_em_timing.log
