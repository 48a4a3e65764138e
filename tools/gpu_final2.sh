#!/bin/bash
# round-end style run: smoke, bench (both arms), C2 probe, ncu launch list + one full capture (window regime)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/bench_c3.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference.json 2>/dev/null; echo "ref rc=$?"
cut -c1-400 gpurun_out/bench_c3_reference.json; echo
timeout 300 python tools/perf_probe.py --N 100000 --D 2 --K 30 --power 1.0 --sweeps 6 > gpurun_out/probe_c2.log 2>&1
grep -o "^sweep [0-9]*\|'K': [0-9]*\|'moves': [0-9]*\|'sweep_kernel_ms': [0-9.]*" gpurun_out/probe_c2.log | paste - - - -
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep \
    python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
