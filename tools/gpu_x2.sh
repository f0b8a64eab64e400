#!/bin/bash
# xprop2 iteration: parity of every variant, then per-op timings for the old kernel (BSMM_XPROP2=0), the automatic choice and
# each forced variant at BASELINE cfg 2.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py -q -k "xprop2" > gpurun_out/pytest_x2.txt 2>&1; rc=$?; tail -3 gpurun_out/pytest_x2.txt
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)" gpurun_out/pytest_x2.txt | head -20; fi
for v in 0 -1 1 2 3; do
  echo "== BSMM_XPROP2=$v" | tee -a gpurun_out/time_x2.txt
  BSMM_XPROP2=$v timeout 300 python tools/time_ops.py ${DENS:-0.05 0.10 0.25 0.5} 2>&1 | sed 's/updat.*//' | tee -a gpurun_out/time_x2.txt
done
