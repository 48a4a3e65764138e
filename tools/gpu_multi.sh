#!/bin/bash
# N-GPU bench lines (torchrun, one rank per GPU): c3 independent chains, c5 disjoint shards, c4 independent chains
N=${1:-2}
mkdir -p gpurun_out
for wl in ${WLS:-c3 c5 c4}; do
  steps=6; [ $wl = c4 ] && steps=4
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
     bench.py --gpus $N --workload $wl --steps $steps --warmup 3 > gpurun_out/bench_${wl}_${N}gpu.json 2> gpurun_out/bench_${wl}_${N}gpu.err
  echo "$wl x$N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_${N}gpu.json"))
    print("${wl}", "n_gpus", d["n_gpus"], "value %.3e e2e %.3e ms/step %.1f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "per_rank ms", d.get("per_rank",{}).get("ms"), "gather", d.get("gather_assignments_ms"))
except Exception as e: print("${wl}: no line", e)
PY
  grep -c "NCCL INFO.*Init COMPLETE\|NCCL INFO comm" gpurun_out/bench_${wl}_${N}gpu.err
done
