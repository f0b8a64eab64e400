#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, ncu capture.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 200 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; echo "pytest_tc rc=$rc" >> gpurun_out/pytest_tc.txt
tail -25 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then echo "TC tests failed: stopping early"; exit 1; fi
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 --sweep > gpurun_out/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_xprop -s 2 -c 1 -f -o gpurun_out/xprop python tools/run_xprop.py 0.25 2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu.log
tail -15 gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/smoke.txt; tail -3 gpurun_out/ncu.log; tail -c 2500 gpurun_out/bench.txt
