mkdir -p gpurun_out
python - <<'PY'
import os, sys
sys.path.insert(0, 'tests')
import rsem_files as rf
base = rf.gen_dataset('/tmp/gb', read_type=1, M=50000, N1=1000000, N0=50000, avg_family=10, read_len=100, seed=11)
rf.run_em(base, 1, 'ref', rounds=12, threads=os.cpu_count())
PY
cd /tmp/gb
RSEM_B200_TIMING=1 ncu --set full --clock-control none --import-source on -k regex:gibbs_parallel -c 1 -o $GRAFT_REPO_ROOT/gpurun_out/r19_gibbs $GRAFT_REPO_ROOT/bin/rsem-run-gibbs ref/r s.temp/s s.stat/s 12 2 1 -p 1 --seed 5 -q > $GRAFT_REPO_ROOT/gpurun_out/r19_ncu.log 2>&1
tail -3 $GRAFT_REPO_ROOT/gpurun_out/r19_ncu.log
