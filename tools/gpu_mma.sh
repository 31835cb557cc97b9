#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_flow_gpu.py -q -m gpu -s -x -k "mma_" > gpurun_out/mma.log 2>&1; echo "exit: $?" >> gpurun_out/mma.log
grep -E "^E  |FAILED|passed|failed|exit|rror" gpurun_out/mma.log | tail -8
timeout 150 python tools/bench_kernels.py --only flow 2>&1 | grep -E "UF|TF-1D|TF-A"
