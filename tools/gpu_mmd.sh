#!/bin/bash
# MMD loss: parity tests, timing next to the reference's torch implementation, launch list, then the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_mmd_gpu.py -q -m gpu > gpurun_out/mmd_tests.log 2>&1; echo "exit: $?" >> gpurun_out/mmd_tests.log
tail -15 gpurun_out/mmd_tests.log
timeout 100 python tools/bench_mmd.py > gpurun_out/mmd_bench.log 2>&1; echo "exit: $?" >> gpurun_out/mmd_bench.log
cat gpurun_out/mmd_bench.log
timeout 60 python tools/bench_mmd.py --batch 1024 --dim 12 > gpurun_out/mmd_bench_1024.log 2>&1; cat gpurun_out/mmd_bench_1024.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/mmd_launches.csv \
  python tools/bench_mmd.py --batch 4096 --dim 10 > gpurun_out/mmd_ncu.log 2>&1
grep -c mmd_pair gpurun_out/mmd_launches.csv
[ "$MMD_FULL" = 1 ] && timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/gpu_tests_mmd_round.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests_mmd_round.log
tail -3 gpurun_out/gpu_tests_mmd_round.log
