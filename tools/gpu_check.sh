# gpurun --timeout 600 -- 'bash tools/gpu_check.sh': GPU parity tests, a short bench and the per-phase cycle profile into gpurun_out/check/
set -x
D=gpurun_out/check; mkdir -p $D
python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
python bench.py --steps 5 --warmup 3 --frames 1184 --e2e-frames 64 --no-cpu-baseline > $D/bench.log 2>&1
python tools/phase_profile.py > $D/phase.log 2>&1
