#!/bin/bash
# configs 3 and 5 at the sizes BASELINE.json states (1e7 events; 1e4 events x 1e4 points) + the context-alignment test
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_flow_gpu.py -q -m gpu -x -k "any_alignment" > gpurun_out/tests_align.log 2>&1; tail -3 gpurun_out/tests_align.log
timeout 600 python tools/bench_configs.py --configs 3,5 --events3 10000000 --events5 10000 --points5 10000 --steps 1 > gpurun_out/configs_stated.jsonl 2>&1; tail -3 gpurun_out/configs_stated.jsonl
