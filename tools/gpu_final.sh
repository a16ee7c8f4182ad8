# gpurun --timeout 2400 -- 'bash tools/gpu_final.sh': the round's official local run into gpurun_out/final/: full GPU suite (default mode and
# software clusters only), smoke(), the bench lines (C4, reference arm, C3 with labels, latency), per-phase profile, per-preset / per-mode /
# slow-path rates, the C5 sweep, the launch lists, one ncu --set full capture, and the in-situ phase costs (phase-doubling builds in
# groundsegmentationrosnode_b200/variants/, see GS_EXP_DOUBLE in csrc/gs_fused.cu).  Numbers taken under ncu are never bench values.
set -x
D=gpurun_out/final; mkdir -p $D
python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
GS_FUSED_MODE=soft timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $D/pytest_soft.log 2>&1; echo "pytest rc=$?" >> $D/pytest_soft.log
python __graft_entry__.py --smoke > $D/smoke.log 2>&1; echo "smoke rc=$?" >> $D/smoke.log
python bench.py > $D/bench.json 2> $D/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
python bench.py --config C3 --no-cpu-baseline > $D/bench_c3.json 2> $D/bench_c3.err
python bench.py --latency > $D/latency.json 2> $D/latency.err
python tools/phase_profile.py > $D/phase.log 2>&1
python tools/preset_throughput.py > $D/presets.log 2>&1
python tools/slow_paths.py > $D/slow_paths.txt 2>&1
python tools/sweep_c5.py > $D/sweep_c5.txt 2>&1
for m in hw soft hybrid; do
  GS_FUSED_MODE=$m timeout 300 python bench.py --steps 5 --warmup 3 --frames 2048 --e2e-frames 64 --no-cpu-baseline --no-parity > $D/bench_mode_$m.log 2> $D/bench_mode_$m.err
done
python tools/ccl_bench.py 3 256 > $D/ccl.log 2>&1
python tools/ccl_bench.py 2 512 >> $D/ccl.log 2>&1
python tools/ccl_bench.py 3 64 > $D/ccl_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $D/ccl_launches.csv python tools/ccl_bench.py 3 64 > $D/ccl_ncu.log 2>&1
bash tools/ab_bench.sh > $D/ab.log 2>&1; cp gpurun_out/ab/ab.txt $D/phase_cost_ab.txt
python bench.py --steps 2 --warmup 3 --frames 592 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 3 --frames 592 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/launches.log 2>&1
python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 2 -o $D/prof_r2 -f python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/ncu.log 2>&1
