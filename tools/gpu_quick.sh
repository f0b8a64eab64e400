#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; echo "pytest_tc rc=$rc" >> gpurun_out/pytest_tc.txt
tail -5 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then echo "TC tests failed: stopping early"; tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
echo "--- OCC=2"; BSMM_XPROP_OCC=2 timeout 300 python tools/time_ops.py 2>&1 | tee gpurun_out/time_occ2.txt
echo "--- OCC=1"; BSMM_XPROP_OCC=1 timeout 300 python tools/time_ops.py 2>&1 | tee gpurun_out/time_occ1.txt
