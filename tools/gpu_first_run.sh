#!/bin/bash
# first GPU contact: RAMBO parity tests (TMA path, then plain staging path)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests/test_rambo_gpu.py -x -q -m gpu -s > gpurun_out/rambo_tests.log 2>&1
echo "exit: $?" >> gpurun_out/rambo_tests.log
tail -30 gpurun_out/rambo_tests.log
MF_NO_TMA=1 python -m pytest tests/test_rambo_gpu.py -x -q -m gpu > gpurun_out/rambo_tests_notma.log 2>&1
echo "exit: $?" >> gpurun_out/rambo_tests_notma.log
tail -5 gpurun_out/rambo_tests_notma.log
