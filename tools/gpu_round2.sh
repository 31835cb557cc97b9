#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): GPU tests, smoke, bench (with the CPU arm), kernel / config timings, the ncu
# launch list of the bench command and one `ncu --set full` capture of the flow kernel, probes, per-warp trace.
# Every step is bounded by its own timeout; outputs land in gpurun_out/ (scratch) and are summarised into profiles/.
set -x
mkdir -p gpurun_out
timeout 90 python tools/gate_tc.py > gpurun_out/gate_final.log 2>&1 || { echo "GATE FAILED"; tail -5 gpurun_out/gate_final.log; exit 7; }
timeout 700 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "exit: $?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "exit: $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; tail -c 600 gpurun_out/bench_full.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
timeout 400 python tools/bench_kernels.py > gpurun_out/kernels.txt 2>&1; cat gpurun_out/kernels.txt
timeout 300 python tools/bench_configs.py > gpurun_out/configs.jsonl 2>&1; tail -3 gpurun_out/configs.jsonl
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 100 python tools/profile_tc.py > gpurun_out/prof_tc_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc python tools/profile_tc.py > gpurun_out/ncu_tc.log 2>&1
timeout 60 python tools/probe_tmem.py > gpurun_out/probe_tmem.txt 2>&1
timeout 60 python tools/probe_spline.py > gpurun_out/probe_spline.txt 2>&1
timeout 300 python tools/bench_train.py --steps 3 --warmup 2 > gpurun_out/train.jsonl 2>&1
timeout 200 python tools/bench_train.py --steps 3 --warmup 2 --no-sampling-loss >> gpurun_out/train.jsonl 2>&1
ls -la gpurun_out | tail -25
