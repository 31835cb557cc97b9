#!/bin/bash
# One optimisation iteration on a B200 (gpurun): hang gate, tcgen05 parity tests, bench, per-warp trace.
# Every step is bounded by its own timeout; a kernel that hangs fails the 40 s gate before anything longer starts.
set -x
mkdir -p gpurun_out
TAG=${1:-iter}
MODE=${2:-all}
timeout 60 python tools/gate_tc.py > gpurun_out/gate_$TAG.log 2>&1 || { echo "GATE FAILED"; tail -5 gpurun_out/gate_$TAG.log; exit 7; }
if [ "$MODE" = all ]; then
timeout 500 python -m pytest tests/test_flow_gpu.py -q -m gpu -x -k "tc or benchmark_scale or group" > gpurun_out/tests_$TAG.log 2>&1; echo "exit: $?" >> gpurun_out/tests_$TAG.log
tail -3 gpurun_out/tests_$TAG.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -c 700 gpurun_out/bench_$TAG.log
timeout 200 python tools/bench_kernels.py --only flow > gpurun_out/kernels_$TAG.txt 2>&1; tail -12 gpurun_out/kernels_$TAG.txt
fi
MF_TC_TRACE=1 timeout 300 python -c "from memflow_msci_project_b200 import build; build.build(force=True)" > gpurun_out/build_trace.log 2>&1
timeout 60 python tools/trace_tc.py tc_3xf16 2600 > gpurun_out/trace_$TAG.txt 2>&1
tail -2 gpurun_out/trace_$TAG.txt
for EXP in $3; do   # extra traced builds with experiment macros
  MF_TC_TRACE=1 MF_TC_DEFS=$EXP timeout 300 python -c "from memflow_msci_project_b200 import build; build.build(force=True)" > gpurun_out/build_trace_$EXP.log 2>&1
  timeout 60 python tools/trace_tc.py tc_3xf16 2600 > gpurun_out/trace_${TAG}_$EXP.txt 2>&1
done
