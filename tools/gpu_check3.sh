#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multichain.py tests/test_gpu_sweep_parity.py tests/test_gpu_large.py tests/test_gpu_integration_stub.py -m gpu -q --timeout 300 --timeout-method=thread -x > gpurun_out/pytest_new.log 2>&1
echo "subset tests rc=$?" | tee -a gpurun_out/pytest_new.log
tail -12 gpurun_out/pytest_new.log
timeout 600 python bench.py --chains 64 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
python - <<'PY'
import json
for wl in ("c3",):
    d=json.load(open("gpurun_out/bench_%s.json"%wl))
    print(wl, "value %.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), "ms", d["config"]["ms_per_sweep"], "moves", d["config"]["moves_per_sweep"])
    print(" warm", d["warmup_chain"]["ms"], "multi", d.get("multi_chain"))
PY
timeout 300 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-200 gpurun_out/bench_c2.json
timeout 400 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-200 gpurun_out/bench_c5.json
# phase clocks (profile build)
export BGMM_B200_LIB=$PWD/pybgmm_b200/lib/libbgmm_b200_prof.so
timeout 200 python tools/perf_probe.py --N 200000 --D 16 --K 100 --sweeps 3 > gpurun_out/probe_c3_prof.log 2>&1
timeout 200 python tools/perf_probe.py --N 100000 --D 2 --K 30 --power 1.0 --sweeps 2 > gpurun_out/probe_c2_prof.log 2>&1
grep "phases\|sweep " gpurun_out/probe_c3_prof.log gpurun_out/probe_c2_prof.log | cut -c1-600
