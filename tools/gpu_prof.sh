#!/bin/bash
mkdir -p gpurun_out
for nz in 1 2 3; do
echo "--- BGMM_NEAR=$nz"
BGMM_NEAR=$nz timeout 300 python tools/perf_probe.py --sweeps 5 > gpurun_out/probe_n$nz.log 2>&1
grep -o "^sweep [0-9]*\|'moves': [0-9]*\|'windows': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe_n$nz.log | paste - - - - | tail -4
done
BGMM_NEAR=2 BGMM_WPROF=1 BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe_prof.log 2>&1
grep "phases\|unit" gpurun_out/probe_prof.log | cut -c1-420 | tail -13
