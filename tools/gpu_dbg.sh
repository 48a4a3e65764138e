#!/bin/bash
mkdir -p gpurun_out
for T in 16 32 0; do
  echo "=== BGMM_TUNE=$T"
  BGMM_TUNE=$T timeout 300 python -m pytest tests/test_gpu_multichain.py::test_forked_chain_equals_a_chain_of_its_own "tests/test_gpu_sweep_parity.py::test_crp_sweeps_match_oracle" tests/test_gpu_large.py::test_replicas_are_deterministic -m gpu -q --timeout 200 --timeout-method=thread 2>&1 | tail -4 | cut -c1-200
done
BGMM_TUNE=16 timeout 300 python bench.py --no-cpu > gpurun_out/bench_c3_red.json 2> gpurun_out/bench_c3_red.err; echo "bench(REDs) rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_c3_red.json"))
    print("c3 REDs value %.3e"%d["value"], "ms", d["config"]["ms_per_sweep"], "warm", d["warmup_chain"]["ms"])
except Exception as e: print("no bench", e)
PY
