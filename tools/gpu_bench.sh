#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
cat gpurun_out/bench_c3.json | cut -c1-3000
tail -3 gpurun_out/bench_c3.err
