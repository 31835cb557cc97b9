#!/bin/bash
# round-1d evidence: kernel timings, configs, wide-flow kernel profile, legacy-MMA probe
mkdir -p gpurun_out
timeout 400 python tools/bench_kernels.py > gpurun_out/kernels.txt 2>&1; tail -30 gpurun_out/kernels.txt
timeout 300 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; tail -4 gpurun_out/configs.jsonl
timeout 100 python tools/probe_mma_sync.py > gpurun_out/probe_mma_sync.txt 2>&1
timeout 60 python tools/profile_mma.py > gpurun_out/prof_mma_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flow_mma_kernel -s 1 -c 1 -o gpurun_out/prof_mma python tools/profile_mma.py > gpurun_out/ncu_mma.log 2>&1
tail -1 gpurun_out/ncu_mma.log
