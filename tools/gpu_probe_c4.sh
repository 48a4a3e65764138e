#!/bin/bash
# D = 64: cluster engine + window kernel parity tests, then the C4 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cluster.py "tests/test_gpu_fullsize.py::test_benchmarked_chain_matches_oracle" -x -q --timeout 600 --timeout-method=thread > gpurun_out/pytest_bigwin.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_bigwin.log | cut -c1-300
timeout 600 python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c4.json"))
print("c4 value %.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), d["config"]["ms_per_sweep"], d["config"]["moves_per_sweep"], "launches", d["gpu_launches"], "roofline", d["roofline"]["bound"], d["roofline"]["frac"], "sweep0", d["sweep0"])
PY
