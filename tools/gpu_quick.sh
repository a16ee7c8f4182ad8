# gpurun --timeout 900 -- 'bash tools/gpu_quick.sh': tools being tuned
set -x
D=gpurun_out/quick; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "labels or label_components" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
timeout 200 python tools/ccl_bench.py 3 256 > $D/ccl.log 2>&1
timeout 200 python tools/ccl_bench.py 2 512 >> $D/ccl.log 2>&1
timeout 300 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > $D/bench_c3.log 2> $D/bench_c3.err
