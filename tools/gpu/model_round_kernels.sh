mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -8 > gpurun_out/r45_tests_dropin.log
cat gpurun_out/r45_tests_dropin.log
python - <<'PY'
import sys
sys.path.insert(0, 'tests')
import rsem_files as rf
rf.gen_dataset('/tmp/gm', read_type=3, M=50000, N1=1000000, N0=50000, avg_family=10, read_len=100, seed=11)
PY
cd /tmp/gm
export RSEM_MAX_ROUND=13 RSEM_MIN_ROUND=13
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conprb_kernel|update_|fold_replicas" --csv --log-file $GRAFT_REPO_ROOT/gpurun_out/r45_launches_model.csv $GRAFT_REPO_ROOT/bin/rsem-run-em ref/r 3 s s.temp/s s.stat/s -p 32 -q > $GRAFT_REPO_ROOT/gpurun_out/r45_ncu.log 2>&1
python - <<'PY'
import csv,collections,os
rows=[r for r in csv.reader(open(os.environ['GRAFT_REPO_ROOT']+'/gpurun_out/r45_launches_model.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0].replace('void ','').replace('unnamed>::','')[:60]
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[-1])/1e6
for k,v in agg.items(): print('%-62s n=%3d total %.3f ms avg %.4f ms'%(k,v[0],v[1],v[1]/v[0]))
PY
