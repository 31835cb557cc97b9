#!/bin/bash
# final tree: hang gate, the whole GPU suite, smoke, bench (both arms)
set -x
mkdir -p gpurun_out
timeout 90 python tools/gate_tc.py > gpurun_out/gate_final.log 2>&1 || { echo "GATE FAILED"; tail -5 gpurun_out/gate_final.log; exit 7; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests_final.log
tail -4 gpurun_out/gpu_tests_final.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "exit: $?" >> gpurun_out/smoke_final.log; tail -2 gpurun_out/smoke_final.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2>&1; tail -c 400 gpurun_out/bench_final.log
