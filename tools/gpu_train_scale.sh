#!/bin/bash
# training step (likelihood) under torchrun on N GPUs: weak scaling, all-reduce share
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss > gpurun_out/train_n${N}.log 2>&1; tail -c 700 gpurun_out/train_n${N}.log
