# gpurun --timeout 900 -- 'bash tools/gpu_quick.sh': tools being tuned
set -x
D=gpurun_out/quick; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_api.py -m gpu -x -q -k "repair or full_pipeline or random_sensor or host_api" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
timeout 300 python tools/preset_throughput.py > $D/presets.log 2>&1
