# K2 on matrices of other shapes: mean hits per read 2 .. 60 (the lanes-per-row choice follows the mean degree)
mkdir -p gpurun_out
for D in 2 4 8 14 30 60; do
S=$(python -c "print(min(1.0, 6.0/$D))")
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 --scale $S --deg $D > gpurun_out/deg_$D.log 2>&1
echo "deg=$D scale=$S"; grep '^{' gpurun_out/deg_$D.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['hits_per_gpu'], d['value'], d['roofline']['k2_ms_per_launch'], d['roofline']['frac'])"
done
