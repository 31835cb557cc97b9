#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_flow_gpu.py -q -m gpu -s -k "hypernet or masked_linear or autograd or gradients" > gpurun_out/tests_autograd.log 2>&1; grep "passed\|failed\|Error" gpurun_out/tests_autograd.log | tail -4 | cut -c1-300
