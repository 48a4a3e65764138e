#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep \
    python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 4 > gpurun_out/ncu_full.log 2>&1
BGMM_WPROF=1 BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so timeout 300 python tools/perf_probe.py --sweeps 4 > gpurun_out/probe_prof.log 2>&1
ls -la gpurun_out/prof_sweep.ncu-rep
