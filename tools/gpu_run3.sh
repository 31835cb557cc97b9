#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests.log
grep -E "max err|fp32 kernel|mismatch|^E  |FAILED|passed|failed|exit" gpurun_out/gpu_tests.log | tail -70
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit: $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "exit: $?" >> gpurun_out/bench.log; tail -5 gpurun_out/bench.log
