D=gpurun_out/s22; mkdir -p $D
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitizer_case.py > $D/memcheck.log 2>&1; echo "rc=$?" >> $D/memcheck.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 0 python tools/sanitizer_case.py > $D/racecheck.log 2>&1; echo "rc=$?" >> $D/racecheck.log
