#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench, C4, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > $O/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -n 16 $O/r2z_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/r2z_smoke.log
timeout 400 python bench.py > $O/r2z_bench_default.log 2>&1; echo "bench rc=$?"
tail -n 1 $O/r2z_bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','clocks')})[:1800]); print(json.dumps(d['roofline'])[:700])"
RSEM_B200_TIMING=1 timeout 300 python bench.py --workload C4 > $O/r2z_bench_C4.log 2>&1; echo "C4 rc=$?"; grep "gibbs kernels" $O/r2z_bench_C4.log | tail -2; tail -n 1 $O/r2z_bench_C4.log | cut -c1-400
