#!/bin/bash
# deliverable run: regression, bench line, ncu launch list of the bench command, one full capture of the sweep kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep \
    python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 4 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
cut -c1-1500 gpurun_out/bench_c3.json
