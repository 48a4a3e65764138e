#!/bin/bash
# first GPU pass: parity tests, engine probe at C3 size, default bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 6 > gpurun_out/probe_c3.log 2>&1
tail -8 gpurun_out/probe_c3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_c3.json
