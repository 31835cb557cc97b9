#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_rambo_gpu.py -x -q -m gpu > gpurun_out/rambo_tests.log 2>&1; echo "exit: $?" >> gpurun_out/rambo_tests.log
tail -3 gpurun_out/rambo_tests.log
python -m pytest tests/test_flow_gpu.py -q -m gpu -s > gpurun_out/flow_tests.log 2>&1; echo "exit: $?" >> gpurun_out/flow_tests.log
tail -60 gpurun_out/flow_tests.log
