mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, 'tests')
import rsem_files as rf
rf.gen_dataset('/tmp/gb', read_type=1, M=50000, N1=2000000, N0=100000, avg_family=10, read_len=100, seed=11)
PY
cd /tmp/gb
export RSEM_MAX_ROUND=20 RSEM_MIN_ROUND=20 RSEM_B200_TIMING=1
( time $GRAFT_REPO_ROOT/bin/rsem-run-em ref/r 1 s s.temp/s s.stat/s -p 32 --gibbs-out -q ) > $GRAFT_REPO_ROOT/gpurun_out/r24_em_timing.log 2>&1
( time $GRAFT_REPO_ROOT/bin/rsem-run-gibbs ref/r s.temp/s s.stat/s 50 64 1 -p 8 --seed 3 -q ) >> $GRAFT_REPO_ROOT/gpurun_out/r24_em_timing.log 2>&1
$GRAFT_REPO_ROOT/oracle/_ref/rsem-build-read-index 32 1 1 s.temp/s_alignable.fq
( time $GRAFT_REPO_ROOT/oracle/_ref/rsem-run-em-rounds ref/r 1 s2 s.temp/s s.stat/s -p 128 --gibbs-out -q ) >> $GRAFT_REPO_ROOT/gpurun_out/r24_em_timing.log 2>&1
cat $GRAFT_REPO_ROOT/gpurun_out/r24_em_timing.log
