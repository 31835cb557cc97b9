#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/tests_gemm.log 2>&1; tail -12 gpurun_out/tests_gemm.log
timeout 300 python -m pytest tests/test_flow_gpu.py -q -m gpu -x -s -k "hypernet or masked_linear or autograd or gradients" > gpurun_out/tests_autograd.log 2>&1; grep "passed\|failed\|Error" gpurun_out/tests_autograd.log | tail -4 | cut -c1-300
timeout 120 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cat gpurun_out/bench_gemm.txt
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss > gpurun_out/train_native.jsonl 2>&1; cut -c1-260 gpurun_out/train_native.jsonl
