#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_flow_gpu.py -q -m gpu -s -x -k "tc_" > gpurun_out/tc.log 2>&1; echo "exit: $?" >> gpurun_out/tc.log
grep -E "TC |kernel-fp32|mismatch|^E  |FAILED|passed|failed|exit|rror" gpurun_out/tc.log | tail -60
