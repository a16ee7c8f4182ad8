D=gpurun_out/s35; mkdir -p $D
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "union" > $D/pytest.log 2>&1; echo "rc=$?" >> $D/pytest.log
