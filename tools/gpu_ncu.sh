#!/bin/bash
# ncu evidence, summarised on the box (the reports are too large to bring back)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 300 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_|softmax|sdd|dds" -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 6 -c 3 -f -o /tmp/r2_tc_kernels python tools/run_xprop.py 0.25 3 > gpurun_out/ncu.log 2>&1; echo "ncu tc rc=$?"
python tools/ncu_summary.py /tmp/r2_tc_kernels.ncu-rep > gpurun_out/r2_ncu_tc_kernels.txt 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:softmax -s 4 -c 2 -f -o /tmp/r2_softmax python tools/bench_bst.py > gpurun_out/ncu2.log 2>&1; echo "ncu softmax rc=$?"
python tools/ncu_summary.py /tmp/r2_softmax.ncu-rep > gpurun_out/r2_ncu_softmax.txt 2>&1
BS=16 timeout 500 ncu --set full --clock-control none -k regex:tc_ -s 6 -c 3 -f -o /tmp/r2_bs16 python tools/time_ops.py 0.2 > gpurun_out/ncu3.log 2>&1; echo "ncu bs16 rc=$?"
python tools/ncu_summary.py /tmp/r2_bs16.ncu-rep 2>&1 | head -120 > gpurun_out/r2_ncu_bs16.txt
timeout 200 python tools/host_cost.py > gpurun_out/host_cost.txt 2>&1; cat gpurun_out/host_cost.txt
du -sh gpurun_out; head -50 gpurun_out/r2_ncu_tc_kernels.txt
