#!/bin/bash
# what the driver runs at round end on one GPU: the GPU test-suite, smoke(), the reference arm and bench.py
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.txt | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.txt 2>&1; cut -c1-160 gpurun_out/bench_ref.txt | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.txt
timeout 100 python tools/host_cost.py | tee gpurun_out/host_cost.txt
