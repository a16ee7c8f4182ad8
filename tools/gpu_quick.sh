set -x
D=gpurun_out/quick; mkdir -p $D
python bench.py --latency --no-cpu-baseline > $D/latency.log 2> $D/latency.err
python bench.py --latency --no-cpu-baseline --latency-frames 600 > $D/latency2.log 2> $D/latency2.err
