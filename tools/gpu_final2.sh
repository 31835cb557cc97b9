#!/bin/bash
# final tree of round 2: ncu --set full of the MMD pair kernel, then hang gate, the whole GPU suite, smoke, bench
set -x
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:mmd_pair -s 12 -c 1 -o gpurun_out/ncu_full_mmd_r2 -f \
  python tools/bench_mmd.py --batch 4096 --dim 10 > gpurun_out/ncu_full_mmd.log 2>&1; echo "ncu exit: $?"
timeout 90 python tools/gate_tc.py > gpurun_out/gate_final.log 2>&1 || { echo "GATE FAILED"; tail -5 gpurun_out/gate_final.log; exit 7; }
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests_final.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests_final.log
tail -4 gpurun_out/gpu_tests_final.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "exit: $?" >> gpurun_out/smoke_final.log; tail -2 gpurun_out/smoke_final.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.log 2>&1; tail -c 600 gpurun_out/bench_final.log
