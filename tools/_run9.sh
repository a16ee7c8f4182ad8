D=gpurun_out/s26; mkdir -p $D
python __graft_entry__.py --smoke > $D/smoke.log 2>&1; echo "rc=$?" >> $D/smoke.log
python -m pytest tests -m gpu -q > $D/pytest.log 2>&1; echo "rc=$?" >> $D/pytest.log
