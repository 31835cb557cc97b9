#!/bin/bash
# native hyper-network backward: GEMM tests, autograd tests, GEMM and training-step timings (native vs torch GEMMs)
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_gpu.py -q -m gpu > gpurun_out/tests_gemm.log 2>&1; tail -5 gpurun_out/tests_gemm.log
timeout 300 python -m pytest tests/test_flow_gpu.py -q -m gpu -s -k "hypernet or masked_linear or autograd or gradients" > gpurun_out/tests_autograd.log 2>&1; grep "passed\|failed\|Error" gpurun_out/tests_autograd.log | tail -6 | cut -c1-300
timeout 120 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cat gpurun_out/bench_gemm.txt
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss > gpurun_out/train_native.jsonl 2>&1
MEMFLOW_B200_TORCH_GEMM=1 timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss >> gpurun_out/train_native.jsonl 2>&1
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 >> gpurun_out/train_native.jsonl 2>&1
MEMFLOW_B200_TORCH_GEMM=1 timeout 300 python tools/bench_train.py --steps 3 --warmup 2 >> gpurun_out/train_native.jsonl 2>&1
cut -c1-300 gpurun_out/train_native.jsonl
