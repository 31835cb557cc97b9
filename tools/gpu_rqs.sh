#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_flow_gpu.py -q -m gpu -x -k "standalone or tc_" > gpurun_out/rqs_tests.log 2>&1; echo "exit: $?" >> gpurun_out/rqs_tests.log
grep -E "^E  |FAILED|passed|failed|exit" gpurun_out/rqs_tests.log | tail -8
timeout 200 python tools/bench_kernels.py --only rqs 2>&1 | tail -3
