#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_transformer_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; tail -2 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
timeout 300 python tools/bench_bst.py 2>&1 | cut -c1-150 | tee gpurun_out/bench_bst.txt
