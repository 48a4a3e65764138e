#!/bin/bash
mkdir -p gpurun_out
export BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so
BGMM_WPROF=1 timeout 200 python tools/perf_probe.py --N 200000 --D 16 --K 100 --sweeps 1 > gpurun_out/probe_clu_c3_tl.log 2>&1
BGMM_WPROF=1 timeout 200 python tools/perf_probe.py --N 100000 --D 2 --K 30 --sweeps 1 --power 1.0 > gpurun_out/probe_clu_c2_tl.log 2>&1
grep "step timeline\|cluster step, prep" gpurun_out/probe_clu_c3_tl.log gpurun_out/probe_clu_c2_tl.log | cut -c1-700
