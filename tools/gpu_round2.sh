#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 60 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 60 python tools/perf_probe.py --N 100000 --D 16 --K 100 --sweeps 6 > gpurun_out/probe_1e5.log 2>&1
tail -8 gpurun_out/probe_1e5.log
timeout 150 python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 8 > gpurun_out/probe_c3.log 2>&1
tail -10 gpurun_out/probe_c3.log
