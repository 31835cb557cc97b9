#!/bin/bash
# 8-GPU runs: headline bench (weak scaling), config 3 (10M events sharded), training step with NCCL all-reduce
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "exit: $?" >> gpurun_out/bench_n$N.log; tail -2 gpurun_out/bench_n$N.log
timeout 300 $TR tools/bench_configs.py --configs 3,5 --events3 10000000 --events5 8000 > gpurun_out/configs_n$N.jsonl 2>&1; tail -2 gpurun_out/configs_n$N.jsonl
timeout 300 $TR tools/bench_train.py --steps 3 --warmup 2 > gpurun_out/train_n$N.jsonl 2>&1; tail -1 gpurun_out/train_n$N.jsonl
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
