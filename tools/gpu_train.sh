#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/tests_gemm.log 2>&1; tail -3 gpurun_out/tests_gemm.log
timeout 300 python -m pytest tests/test_flow_gpu.py -q -m gpu -x -s -k "autograd or hypernet or gradients or grad_requiring" > gpurun_out/tests_autograd.log 2>&1; grep "gradient error\|passed\|failed" gpurun_out/tests_autograd.log | tail -6
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss > gpurun_out/train_native.jsonl 2>&1
MEMFLOW_B200_TORCH_GEMM=1 timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss >> gpurun_out/train_native.jsonl 2>&1
cut -c1-330 gpurun_out/train_native.jsonl
