#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_record.py tests/test_gpu_golden.py -m gpu -x -q --timeout 300 > gpurun_out/pytest_record.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_record.log
tail -15 gpurun_out/pytest_record.log
timeout 600 python tools/record_probe.py > gpurun_out/record_probe.log 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/record_probe.log | cut -c1-1500
