#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_rambo_gpu.py -q -m gpu -s > gpurun_out/rambo_tests.log 2>&1; echo "exit: $?" >> gpurun_out/rambo_tests.log
grep -E "quantiles|^E  |FAILED|passed|failed|exit" gpurun_out/rambo_tests.log | tail -12
timeout 200 python tools/bench_kernels.py --only rambo 2>&1 | tail -9
