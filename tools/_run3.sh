set -x
D=gpurun_out/s12; mkdir -p $D
python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 64 --no-cpu-baseline > $D/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o $D/prof_r1_fused_v11 -f python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 64 --no-cpu-baseline > $D/ncu.log 2>&1
