#!/bin/bash
# Quick GPU iteration: tcgen05 parity tests, then per-op timings at BASELINE cfg 2 and the attention ops of cfg 3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; tail -2 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
timeout 300 python tools/time_ops.py 0.05 0.10 0.25 0.5 1.0 2>&1 | tee gpurun_out/time_ops.txt
timeout 300 python tools/bench_bst.py 2>&1 | cut -c1-200 | tee gpurun_out/bench_bst.txt
