#!/bin/bash
# cluster step engine: parity tests, per-sweep probe at C3 / C2, in-kernel timeline (profile build: tools/build_prof.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_clu.py -x -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_clu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_clu.log | cut -c1-300
timeout 300 python tools/perf_probe.py --N 1000000 --D 16 --K 100 --sweeps 3 2>&1 | tee gpurun_out/probe_clu_c3.log | grep -E "^sweep" | cut -c1-100,290-420
timeout 300 python tools/perf_probe.py --N 100000 --D 2 --K 30 --sweeps 2 --power 1.0 2>&1 | tee gpurun_out/probe_clu_c2.log | grep -E "^sweep" | cut -c1-100,290-420
export BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so
BGMM_WPROF=1 timeout 200 python tools/perf_probe.py --N 200000 --D 16 --K 100 --sweeps 1 > gpurun_out/probe_clu_c3_tl.log 2>&1
BGMM_WPROF=1 timeout 200 python tools/perf_probe.py --N 100000 --D 2 --K 30 --sweeps 1 --power 1.0 > gpurun_out/probe_clu_c2_tl.log 2>&1
grep "step timeline\|cluster step, prep" gpurun_out/probe_clu_c3_tl.log gpurun_out/probe_clu_c2_tl.log | cut -c1-700
