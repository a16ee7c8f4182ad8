set -x
D=gpurun_out/quick; mkdir -p $D
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nearest or policy or full_pipeline or batch_device" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
python tools/slow_paths.py > $D/slow_paths.txt 2>&1
python bench.py --steps 5 --warmup 3 --frames 2048 --no-cpu-baseline > $D/bench.log 2> $D/bench.err
