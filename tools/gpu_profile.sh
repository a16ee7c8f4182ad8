# gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh': the ncu evidence of round 2 into gpurun_out/profile/ (launch list of the bench, one
# ncu --set full capture of the fused kernel), the C5 sweep and the host-API test.  Numbers taken under ncu are never bench values.
set -x
D=gpurun_out/profile; mkdir -p $D
python -m pytest tests/test_host_api.py tests/test_gpu_parity.py -m gpu -x -q -k "host or api_surface or native or foreign" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
python tools/sweep_c5.py > $D/sweep_c5.txt 2>&1
python bench.py --steps 2 --warmup 3 --frames 592 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 3 --frames 592 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/launches.log 2>&1
python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/bench_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_frame -s 6 -c 2 -o $D/prof_r2 -f python bench.py --steps 1 --warmup 3 --frames 296 --e2e-frames 74 --no-cpu-baseline --no-parity > $D/ncu.log 2>&1
