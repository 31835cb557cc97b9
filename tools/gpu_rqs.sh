set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "rqs or spline or standalone" > gpurun_out/tests_rqs.log 2>&1; tail -3 gpurun_out/tests_rqs.log
timeout 200 python tools/bench_kernels.py --only rqs > gpurun_out/kernels_rqs.txt 2>&1; tail -4 gpurun_out/kernels_rqs.txt
