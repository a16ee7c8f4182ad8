# gpurun --timeout 900 -- 'bash tools/gpu_check.sh': GPU parity tests, a short bench and the per-phase cycle profile into gpurun_out/check/
set -x
D=gpurun_out/check; mkdir -p $D
python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
python bench.py --steps 5 --warmup 3 --frames 1184 --e2e-frames 296 > $D/bench.log 2> $D/bench.err
python bench.py --config C3 --steps 5 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline > $D/bench_c3.log 2> $D/bench_c3.err
python bench.py --latency > $D/latency.log 2> $D/latency.err
python bench.py --impl reference --steps 2 --warmup 1 > $D/bench_reference.log 2> $D/bench_reference.err
python tools/phase_profile.py > $D/phase.log 2>&1
