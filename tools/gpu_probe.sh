#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 8 > gpurun_out/probe_c3.log 2>&1
cat gpurun_out/probe_c3.log
