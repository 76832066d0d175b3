mkdir -p gpurun_out
for cfg in "16384 0" "16384 4" "16384 64" "65536 0" "65536 64" "4096 16" "1048576 128"; do
  set -- $cfg
  export RSEM_B200_GIBBS_BLOCK=$1
  if [ "$2" != "0" ]; then export RSEM_B200_GIBBS_CTAS=$2; else unset RSEM_B200_GIBBS_CTAS; fi
  echo "block=$1 ctas=$2" >> gpurun_out/r17_gibbs_sweep.log
  timeout 600 python tools/bench_gibbs.py --N1 1000000 --M 50000 --burnin 60 --nsamples 2 --chains 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['sweeps_per_chain'], d['ours_phases'], d['identical_countvectors'])" >> gpurun_out/r17_gibbs_sweep.log
done
cat gpurun_out/r17_gibbs_sweep.log
