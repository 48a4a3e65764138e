#!/bin/bash
# full-size ncu evidence (one GPU): launch list of a bench run, then --set full of the cold sweep (launch 0) and of the
# window-regime sweep (launch 2) of the C3 chain at N = 1e6
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_bench.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 0 -c 1 -f -o gpurun_out/prof_full_cold \
   python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 1 > gpurun_out/ncu_full_cold.log 2>&1
echo "cold rc=$?"; tail -2 gpurun_out/ncu_full_cold.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 2 -c 1 -f -o gpurun_out/prof_full_window \
   python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 3 > gpurun_out/ncu_full_window.log 2>&1
echo "window rc=$?"; tail -2 gpurun_out/ncu_full_window.log | cut -c1-300
ls -la gpurun_out/prof_full_*.ncu-rep
