# gpurun --timeout 900 -- 'bash tools/gpu_ncu.sh': one ncu --set full capture of the fused kernel (after the same command ran clean without ncu)
set -x
D=gpurun_out/profile; mkdir -p $D
python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 2 -o $D/prof_r2b -f python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/ncu.log 2>&1
