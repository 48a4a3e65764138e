#!/bin/bash
# round-end style run: build check, GPU tests, smoke, bench (both arms), ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 3 -c 1 -f -o gpurun_out/prof_sweep \
    python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 4 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast_sweep -s 7 -c 1 -f -o gpurun_out/prof_sweep_converged \
    python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 8 > gpurun_out/ncu_full2.log 2>&1
cut -c1-700 gpurun_out/bench_c3.json; echo; cut -c1-500 gpurun_out/bench_c3_reference.json
