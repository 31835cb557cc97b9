#!/bin/bash
# full evidence run of the round: tests, smoke, kernel timings, bench (plain), ncu launch list, ncu full profiles
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit: $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 300 python tools/bench_kernels.py > gpurun_out/kernels.txt 2>&1; cat gpurun_out/kernels.txt
timeout 300 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; tail -3 gpurun_out/configs.jsonl
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 100 python tools/profile_tc.py > gpurun_out/prof_tc_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:flow_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc python tools/profile_tc.py > gpurun_out/ncu_tc.log 2>&1
timeout 100 python tools/profile_target.py rambo > gpurun_out/prof_rambo_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"rambo_forward|rambo_inverse" -c 8 -o gpurun_out/prof_rambo python tools/profile_target.py rambo > gpurun_out/ncu_rambo.log 2>&1
timeout 100 python tools/probe_issue.py > gpurun_out/probe_issue.txt 2>&1
timeout 100 python tools/probe_issue_batch.py >> gpurun_out/probe_issue.txt 2>&1
timeout 100 python tools/probe_spline.py > gpurun_out/probe_spline.txt 2>&1
timeout 400 python tools/bench_train.py --steps 3 --warmup 2 > gpurun_out/train.jsonl 2>&1
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss >> gpurun_out/train.jsonl 2>&1
ls -la gpurun_out | tail -20
