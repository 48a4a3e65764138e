#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
grep -o "^sweep [0-9]*\|'moves': [0-9]*\|'windows': [0-9]*\|'seq_data': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe.log | paste - - - - - 
BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe_prof.log 2>&1
grep "phases\|unit" gpurun_out/probe_prof.log | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 0 -c 1 -f -o gpurun_out/prof_sweep_cold \
    python tools/perf_probe.py --N 100000 --D 16 --K 100 --sweeps 1 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/prof_sweep_cold.ncu-rep
