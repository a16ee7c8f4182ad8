D=gpurun_out/s20; mkdir -p $D
B="python bench.py --steps 5 --warmup 3 --frames 1184 --e2e-frames 64 --no-cpu-baseline"
GS_FUSED_THREADS=512 GS_DEBUG_PLAN=1 $B > $D/b_512.log 2>&1
$B > $D/b_1024.log 2>&1
GS_FUSED_THREADS=512 python tools/phase_profile.py > $D/phase512.log 2>&1
