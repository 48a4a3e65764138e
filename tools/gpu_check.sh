#!/bin/bash
# One GPU-box call: the -m gpu suite, then the default bench (C3) and short probes of the other configs.
# usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tests|bench|all]
what=${1:-all}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ "$what" = tests ] || [ "$what" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --durations=8 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
  cut -c1-1500 gpurun_out/bench_c3.json
  timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
  cut -c1-900 gpurun_out/bench_c2.json
  timeout 600 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
  cut -c1-900 gpurun_out/bench_c5.json
fi
