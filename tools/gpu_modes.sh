# gpurun --timeout 1200 -- 'bash tools/gpu_modes.sh': the fused kernel with thread-block clusters only (hw), software clusters only (soft)
# and both side by side (hybrid, the default): parity tests in the soft and the default mode, then the bench in each mode
set -x
D=gpurun_out/modes; mkdir -p $D
GS_FUSED_MODE=soft timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_soft.log 2>&1; echo "pytest rc=$?" >> $D/pytest_soft.log
timeout 600 python -m pytest tests -m gpu -x -q > $D/pytest_default.log 2>&1; echo "pytest rc=$?" >> $D/pytest_default.log
for m in hw soft hybrid; do
  GS_DEBUG_PLAN=1 GS_FUSED_MODE=$m timeout 300 python bench.py --steps 5 --warmup 3 --frames 2048 --e2e-frames 64 --no-cpu-baseline > $D/bench_$m.log 2> $D/bench_$m.err
done
GS_FUSED_MODE=soft timeout 300 python tools/phase_profile.py > $D/phase_soft.log 2>&1
timeout 300 python tools/preset_throughput.py > $D/presets.log 2>&1
