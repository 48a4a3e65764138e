#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_fullsize.py tests/test_gpu_sweep_parity.py -x -q --timeout 600 --timeout-method=thread > gpurun_out/pytest_bigwin.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_bigwin.log | cut -c1-300
timeout 600 python tools/perf_probe.py --N 200000 --D 64 --K 100 --sweeps 5 --power 1.0 2>&1 | tee gpurun_out/probe_bigwin_c4.log | grep -E "^sweep" | cut -c1-130,300-460
