#!/bin/bash
# Milestone check (round 2): full GPU test-suite, smoke, bench (with sub-records), reference arm, host-cost probe,
# ncu launch list of the bench command + full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.txt | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py --steps 300 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.err
python tools/show_bench.py gpurun_out/bench.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_|softmax|sdd|dds" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 6 -c 3 -f -o gpurun_out/r2_tc_kernels python tools/run_xprop.py 0.25 3 > gpurun_out/ncu.log 2>&1; echo "ncu tc rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:softmax -s 4 -c 2 -f -o gpurun_out/r2_softmax python tools/bench_bst.py > gpurun_out/ncu2.log 2>&1; echo "ncu softmax rc=$?"
ls -la gpurun_out/*.ncu-rep
