mkdir -p gpurun_out
(nvidia-smi; nproc; free -g; lscpu | head -20) > gpurun_out/r1_sysinfo.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --scale 0.02 --steps 5 > gpurun_out/r1_bench_small.log 2>&1
python bench.py > gpurun_out/r1_bench_c3.log 2>&1
python bench.py --variant 2 --no-cpu-baseline > gpurun_out/r1_bench_c3_direct.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches.csv python bench.py --scale 0.2 --steps 3 --no-cpu-baseline > gpurun_out/r1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:estep -s 3 -c 2 -o gpurun_out/r1_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
