#!/bin/bash
mkdir -p gpurun_out
# sequential regime: first sweep of a cold chain
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -c 1 -o gpurun_out/prof_seq -f \
   python tools/perf_probe.py --N 30000 --D 16 --K 100 --sweeps 1 > gpurun_out/ncu_seq.log 2>&1
# window regime: third sweep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 2 -c 1 -o gpurun_out/prof_win -f \
   python tools/perf_probe.py --N 100000 --D 16 --K 100 --sweeps 3 > gpurun_out/ncu_win.log 2>&1
tail -3 gpurun_out/ncu_seq.log gpurun_out/ncu_win.log
ls -la gpurun_out/*.ncu-rep
