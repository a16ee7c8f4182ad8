set -x
mkdir -p gpurun_out/s3
python -m pytest tests -m gpu -x -q > gpurun_out/s3/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3/pytest.log
B="python bench.py --steps 5 --warmup 3 --frames 1184 --e2e-frames 64 --no-cpu-baseline"
GS_DEBUG_PLAN=1 $B > gpurun_out/s3/b_default.log 2>&1
GS_FUSED_THREADS=512 GS_DEBUG_PLAN=1 $B > gpurun_out/s3/b_512.log 2>&1
for cl in 4 5 6 7 8; do GS_FUSED_CLUSTER=$cl GS_DEBUG_PLAN=1 $B > gpurun_out/s3/b_cl$cl.log 2>&1; done
python tools/phase_profile.py > gpurun_out/s3/phase.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/s3/smi.log
nproc >> gpurun_out/s3/smi.log
