#!/bin/bash
# e2e path: block cache + class directory built under the conprb copy.  Kernel tests, default bench (phases), the same without
# the cache, launch list of the default bench command restricted to this library's kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python -m pytest tests/test_em_kernels_gpu.py tests/test_capi_library.py -x -q -m gpu > $O/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 $O/r2q_tests.log
RSEM_B200_CLASS_TIMING=1 timeout 400 python bench.py > $O/r2q_bench_default.log 2>&1; echo "bench rc=$?"
grep -E "class layout|phase" $O/r2q_bench_default.log | tail -12
tail -n 1 $O/r2q_bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline')})[:2500])"
RSEM_B200_BLOCK_CACHE=0 timeout 300 python bench.py --no-cpu-baseline --no-traffic > $O/r2q_bench_nocache.log 2>&1; echo "nocache rc=$?"
tail -n 1 $O/r2q_bench_nocache.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d['e2e'])[:900])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"estep|theta_update|cls_|tile_|gather_u|max_d|abs_kernel|RadixSort|DeviceScan|DeviceSelect|reduce_counts" -c 200 --csv --log-file $O/r2q_launches_c3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-traffic > $O/r2q_ncu_launches.log 2>&1; echo "launch list rc=$?"
