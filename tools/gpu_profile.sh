# gpurun --timeout 900 -- 'bash tools/gpu_profile.sh': the default bench, the CPU arm, the ncu launch list and one ncu --set full capture
# of the fused kernel into gpurun_out/profile/ (summaries are made from these with tools/ncu_buckets.py and copied to profiles/)
set -x
D=gpurun_out/profile; mkdir -p $D
python bench.py > $D/bench_default.json 2> $D/bench_default.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 3 --frames 592 --e2e-frames 64 --no-cpu-baseline > $D/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o $D/prof_r1_final -f python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 64 --no-cpu-baseline > $D/ncu.log 2>&1
python tools/phase_profile.py > $D/phase.log 2>&1
