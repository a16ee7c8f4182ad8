# gpurun --timeout 900 -- 'bash tools/gpu_quick.sh': tools being tuned + one compute-sanitizer tool on the small cases
set -x
D=gpurun_out/quick; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "labels or label_components" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
timeout 200 python tools/ccl_bench.py 3 256 > $D/ccl.log 2>&1
timeout 200 python tools/ccl_bench.py 2 512 >> $D/ccl.log 2>&1
timeout 120 python tools/sanitizer_case.py > $D/sanitizer_plain.log 2>&1 &&
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_case.py > $D/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $D/sanitizer_memcheck.log
