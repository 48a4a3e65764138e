#!/bin/bash
# round-2 ncu evidence for the cluster step engine and the D=64 window kernel (one GPU, bounded)
mkdir -p gpurun_out
cap() { # name, skip, kernel regex, probe args...
  name=$1; skip=$2; kre=$3; shift 3
  BGMM_WPROF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/prof_$name \
     python tools/perf_probe.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$?"; grep -c "cluster launch\|window from" gpurun_out/ncu_$name.log
}
cap clu_D16 1 k_clu_sweep --N 1000000 --D 16 --K 100 --sweeps 1
cap clu_D2 1 k_clu_sweep --N 100000 --D 2 --K 30 --sweeps 1 --power 1.0
cap bigwin_D64 14 k_big_window --N 200000 --D 64 --K 100 --sweeps 3 --power 1.0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_bench.csv
ls -la gpurun_out/prof_clu_D16.ncu-rep gpurun_out/prof_clu_D2.ncu-rep gpurun_out/prof_bigwin_D64.ncu-rep
