#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for t in 0 4; do
echo "--- BGMM_TUNE=$t"
BGMM_TUNE=$t timeout 300 python tools/perf_probe.py --sweeps 5 > gpurun_out/probe_t$t.log 2>&1
grep -o "^sweep [0-9]*\|'moves': [0-9]*\|'windows': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe_t$t.log | paste - - - - | tail -4
done
BGMM_WPROF=1 BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe_prof.log 2>&1
grep "phases\|unit" gpurun_out/probe_prof.log | cut -c1-420 | tail -9
