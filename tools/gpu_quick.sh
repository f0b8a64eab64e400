#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q -k "bst" > gpurun_out/pytest_tc.txt 2>&1; rc=$?; echo "pytest_tc rc=$rc" >> gpurun_out/pytest_tc.txt
tail -5 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then echo "TC tests failed: stopping early"; tail -40 gpurun_out/pytest_tc.txt; exit 1; fi
timeout 300 python tools/bench_bst.py 2>&1 | cut -c1-170 | tee gpurun_out/bench_bst.txt
