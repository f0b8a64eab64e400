#!/bin/bash
# One gpurun call: GPU parity tests, tcgen05 probe, smoke, a short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in ; do
  timeout 60 tools/tc_probe $t >> gpurun_out/probe.txt 2>&1 || echo "probe $t rc=$?" >> gpurun_out/probe.txt
done
timeout 120 tools/tc_probe bench >> gpurun_out/probe.txt 2>&1 || echo "probe bench rc=$?" >> gpurun_out/probe.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench.txt
tail -3 gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/smoke.txt; tail -c 600 gpurun_out/bench.txt; tail -50 gpurun_out/probe.txt
