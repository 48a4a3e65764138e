#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multichain.py tests/test_gpu_sweep_parity.py tests/test_gpu_large.py -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/pytest_new.log 2>&1
echo "subset tests rc=$?" | tee -a gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_cluster.py -m gpu -q --timeout 200 --timeout-method=thread > gpurun_out/pytest_cluster.log 2>&1
echo "cluster tests rc=$?" | tee -a gpurun_out/pytest_cluster.log
tail -30 gpurun_out/pytest_cluster.log | cut -c1-250
timeout 600 python bench.py --no-cpu > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for wl in ("c3","c2"):
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%wl))
        print(wl, "value %.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), "ms", d["config"]["ms_per_sweep"], "moves", d["config"]["moves_per_sweep"], "warm", d["warmup_chain"]["ms"])
    except Exception as e: print(wl, "no bench", e)
PY
export BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so
BGMM_WPROF=1 timeout 200 python tools/perf_probe.py --N 100000 --D 16 --K 100 --sweeps 1 > gpurun_out/probe_c3_tl.log 2>&1
grep "timeline\|sweep " gpurun_out/probe_c3_tl.log | cut -c1-330
