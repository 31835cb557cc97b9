#!/bin/bash
# bench.py on N GPUs of one box under torchrun (weak scaling: 1 M events per GPU; strong: 1 M events in total)
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_weak.log 2>&1; tail -c 900 gpurun_out/bench_n${N}_weak.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --scaling strong > gpurun_out/bench_n${N}_strong.log 2>&1; tail -c 900 gpurun_out/bench_n${N}_strong.log
