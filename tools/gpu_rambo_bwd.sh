#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rambo_gpu.py -q -m gpu -x -k "grad or backward" > gpurun_out/tests_rambo_bwd.log 2>&1; tail -5 gpurun_out/tests_rambo_bwd.log
timeout 200 python tools/bench_rambo_bwd.py > gpurun_out/rambo_bwd.txt 2>&1; cat gpurun_out/rambo_bwd.txt
