#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
grep -o "^sweep [0-9]*\|'moves': [0-9]*\|'windows': [0-9]*\|'seq_data': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe.log | paste - - - - - 
BGMM_WPROF=1 BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe_prof.log 2>&1
grep "phases\|unit" gpurun_out/probe_prof.log | cut -c1-420
