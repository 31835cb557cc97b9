#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_probe_gpu.py -q -m gpu -s -x > gpurun_out/probe.log 2>&1; echo "exit: $?" >> gpurun_out/probe.log
grep -E "max err|^E  |FAILED|passed|failed|exit|rror" gpurun_out/probe.log | tail -40
