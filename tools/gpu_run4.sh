#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -s > gpurun_out/gpu_tests.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests.log
grep -E "^E  |FAILED|passed|failed|exit" gpurun_out/gpu_tests.log | tail -30
python tools/diag_flow_fp32.py > gpurun_out/flow_fp32_error.txt 2>&1
python tools/bench_kernels.py > gpurun_out/kernels.txt 2>&1; cat gpurun_out/kernels.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/bench_plain.log
python tools/profile_target.py all > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"flow_kernel|rambo_forward|rambo_inverse" -c 12 -o gpurun_out/prof_r1 python tools/profile_target.py all > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
