D=gpurun_out/s25; mkdir -p $D
B="python bench.py --steps 5 --warmup 3 --frames 1184 --e2e-frames 64 --no-cpu-baseline"
GS_FUSED_THREADS=768 GS_DEBUG_PLAN=1 $B > $D/b_768.log 2>&1
GS_FUSED_THREADS=768 python -m pytest tests/test_gpu_parity.py -q -x -k "full_pipeline or batch_device or union_paths" > $D/pytest768.log 2>&1
GS_FUSED_THREADS=768 python tools/phase_profile.py > $D/phase768.log 2>&1
