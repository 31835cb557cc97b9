#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_flow_gpu.py -q -m gpu -s -x -k "tc_" > gpurun_out/tc.log 2>&1; echo "exit: $?" >> gpurun_out/tc.log
grep -E "^E  |FAILED|passed|failed|exit|rror" gpurun_out/tc.log | tail -6
echo "--- EW=1 (8 epilogue warps)"; timeout 200 python tools/bench_kernels.py --only flow 2>&1 | grep -E "TF-A|TF-B" | grep tc_
echo "--- EW=2 (16 epilogue warps)"; MF_TC_EW=2 timeout 200 python tools/bench_kernels.py --only flow 2>&1 | grep -E "TF-A|TF-B" | grep tc_
