#!/bin/bash
mkdir -p gpurun_out
for wf in 1.0 1.5 3.0; do
echo "--- BGMM_WIN_FACTOR=$wf"
BGMM_WIN_FACTOR=$wf timeout 300 python tools/perf_probe.py --sweeps 5 > gpurun_out/probe_w$wf.log 2>&1
grep -o "^sweep [0-9]*\|'moves': [0-9]*\|'windows': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe_w$wf.log | paste - - - - | tail -4
done
