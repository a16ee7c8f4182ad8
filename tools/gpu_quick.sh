set -x
D=gpurun_out/quick; mkdir -p $D
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nearest or policy or full_pipeline or batch_device or adversarial or shuffl or large_batch or union or random_sensor or lone or beyond" > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
bash tools/ab_bench.sh > $D/ab.log 2>&1
python tools/phase_profile.py > $D/phase.log 2>&1
