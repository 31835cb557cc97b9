#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 100 python tools/profile_gemm.py > gpurun_out/prof_gemm_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 2 -c 2 -o gpurun_out/prof_gemm python tools/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
