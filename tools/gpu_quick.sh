#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; echo "pytest_tc rc=$rc" >> gpurun_out/pytest_tc.txt
tail -5 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then echo "TC tests failed: stopping early"; tail -40 gpurun_out/pytest_tc.txt; exit 1; fi
AXIS=0 timeout 300 python tools/time_ops.py 0.1 0.25 1.0 2>&1 | tee gpurun_out/time_axis0.txt
AXIS=1 BS=64 timeout 300 python tools/time_ops.py 0.1 0.25 1.0 2>&1 | tee gpurun_out/time_bs64.txt
AXIS=0 BS=64 timeout 300 python tools/time_ops.py 0.25 2>&1 | tee -a gpurun_out/time_bs64.txt
