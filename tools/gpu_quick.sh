#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_matmul_gpu.py -x -q -k "gate or golden or gated" > gpurun_out/pytest_tc.txt 2>&1; rc=$?; tail -2 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
