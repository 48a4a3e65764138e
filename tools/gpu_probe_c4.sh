#!/bin/bash
mkdir -p gpurun_out
BGMM_WPROF=1 timeout 300 python tools/perf_probe.py --N 100000 --D 64 --K 100 --power 1.0 --sweeps 2 > gpurun_out/probe_c4.log 2>&1
grep -c "cluster launch" gpurun_out/probe_c4.log
grep "cluster launch" gpurun_out/probe_c4.log | head -12
grep "cluster launch" gpurun_out/probe_c4.log | awk '{print $NF, $(NF-2), $(NF-3)}' | sort | uniq -c | sort -rn | head -5
grep "^sweep\|phases" gpurun_out/probe_c4.log | cut -c1-420
