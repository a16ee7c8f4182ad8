set -x
D=gpurun_out/s16; mkdir -p $D
nvidia-smi -L > $D/gpus.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $D/bench_n2.log 2>&1; echo "rc=$?" >> $D/bench_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $D/ref_n2.log 2>&1; echo "rc=$?" >> $D/ref_n2.log
python bench.py --steps 5 --warmup 3 > $D/bench_n1.log 2>&1; echo "rc=$?" >> $D/bench_n1.log
