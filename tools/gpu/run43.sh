mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:estep_rows -s 3 -c 1 -o gpurun_out/r43_k2 python bench.py --scale 0.2 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r43_ncu.log 2>&1
