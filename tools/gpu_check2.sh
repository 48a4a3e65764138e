#!/bin/bash
# new-engine check: the register-step / multi-chain tests first (short timeouts: a barrier mismatch would hang), then the
# whole -m gpu suite, then the benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multichain.py -m gpu -q --timeout 120 --timeout-method=thread -x > gpurun_out/pytest_new.log 2>&1
echo "new-engine tests rc=$?" | tee -a gpurun_out/pytest_new.log
tail -30 gpurun_out/pytest_new.log
if grep -q "rc=0" gpurun_out/pytest_new.log; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread --durations=8 > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -25 gpurun_out/pytest_gpu.log
  timeout 600 python bench.py --chains 32 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
  cut -c1-1200 gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
  timeout 300 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
  cut -c1-600 gpurun_out/bench_c2.json
  timeout 400 python bench.py --workload c5 --no-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
  cut -c1-600 gpurun_out/bench_c5.json
fi
