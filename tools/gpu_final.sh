# gpurun --timeout 1800 -- 'bash tools/gpu_final.sh': the round's official local run: full GPU suite, smoke(), the bench lines of every mode,
# the per-phase profile, per-preset rates and the launch lists (bench + CCL tool) into gpurun_out/final/
set -x
D=gpurun_out/final; mkdir -p $D
python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
python __graft_entry__.py --smoke > $D/smoke.log 2>&1; echo "smoke rc=$?" >> $D/smoke.log
python bench.py > $D/bench.json 2> $D/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
python bench.py --config C3 --no-cpu-baseline > $D/bench_c3.json 2> $D/bench_c3.err
python bench.py --latency > $D/latency.json 2> $D/latency.err
python tools/phase_profile.py > $D/phase.log 2>&1
python tools/preset_throughput.py > $D/presets.log 2>&1
python tools/ccl_bench.py 3 256 > $D/ccl.log 2>&1
python tools/ccl_bench.py 2 512 >> $D/ccl.log 2>&1
python tools/ccl_bench.py 3 64 > $D/ccl_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $D/ccl_launches.csv python tools/ccl_bench.py 3 64 > $D/ccl_ncu.log 2>&1
