# gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_multi.sh': host-copy ceilings and the native multi-GPU driver at 1/2/4/8 GPUs, then the
# torchrun bench at 8 and 2 ranks (the way the driver launches it)
set -x
D=gpurun_out/multi; mkdir -p $D
nvidia-smi topo -m > $D/topo.txt 2>&1
timeout 500 python tools/e2e_probe.py > $D/e2e_probe.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --frames 1024 --no-cpu-baseline > $D/bench_n8.log 2> $D/bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --frames 1024 --no-cpu-baseline > $D/bench_n2.log 2> $D/bench_n2.err
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native_multi" > $D/pytest_multi.log 2>&1
