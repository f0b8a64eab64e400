#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py -x -q -k "xprop" > gpurun_out/pytest_tc.txt 2>&1; rc=$?; tail -2 gpurun_out/pytest_tc.txt; if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tc.txt; exit 1; fi; timeout 300 python tools/time_ops.py 0.05 0.1 0.25 1.0 2>&1 | tee gpurun_out/time_ops.txt
