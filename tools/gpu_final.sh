#!/bin/bash
# round-2 evidence run (one GPU): the -m gpu suite, smoke(), the bench lines of every configuration, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --chains 128 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c3_s20w5.json 2> gpurun_out/bench_c3_s20w5.err; echo "bench c3 (20,5) rc=$?"
timeout 900 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench_c3_reference.json 2> gpurun_out/bench_c3_reference.err; echo "reference arm rc=$?"
timeout 400 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 600 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
python - <<'PY'
import json
for f in ("bench_c3","bench_c3_s20w5","bench_c3_reference","bench_c2","bench_c4","bench_c5"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "value %.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), "ms/step %.1f"%d["ms_per_step"], (d.get("config") or {}).get("moves_per_sweep"), "frac", (d.get("roofline") or {}).get("frac"), "multi", (d.get("multi_chain") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(f, "no line", e)
PY
