#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_matmul_gpu.py -x -q -k "not updat and not bst" > gpurun_out/pytest_tc.txt 2>&1; rc=$?; tail -2 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
echo "== default"; timeout 300 python tools/time_ops.py 0.05 0.10 0.25 0.5 1.0 2>&1 | cut -c1-150
echo "== bs64"; BS=64 timeout 300 python tools/time_ops.py 0.25 1.0 2>&1 | cut -c1-150
