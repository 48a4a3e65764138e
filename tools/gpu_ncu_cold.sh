#!/bin/bash
# ncu --set full of one cold (sequential-regime) sweep: the first sweep of a small chain from the rand initial state
mkdir -p gpurun_out
N=${1:-60000}; D=${2:-16}; K=${3:-100}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 0 -c 1 -f -o gpurun_out/prof_cold_D$D \
   python tools/perf_probe.py --N $N --D $D --K $K --sweeps 1 $4 > gpurun_out/ncu_cold_D$D.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_cold_D$D.log
ls -la gpurun_out/*.ncu-rep
