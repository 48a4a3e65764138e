#!/bin/bash
# cluster step engine: parity tests, per-sweep probe at C3 / C2 (in-kernel timeline: tools/build_prof.sh + BGMM_WPROF=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clu.py -x -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_clu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_clu.log | cut -c1-300
timeout 300 python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 3 2>&1 | tee gpurun_out/probe_clu_c3.log | grep -E "^sweep" | cut -c1-100,290-420
timeout 300 python tools/perf_probe.py --N 100000 --D 2 --K 30 --sweeps 2 --power 1.0 2>&1 | tee gpurun_out/probe_clu_c2.log | grep -E "^sweep" | cut -c1-100,290-420
