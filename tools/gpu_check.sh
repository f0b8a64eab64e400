#!/bin/bash
# Milestone check: full GPU test-suite, smoke, bench (+sweep), bst bench, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py --sweep --steps 300 > gpurun_out/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1
timeout 300 python tools/bench_bst.py > gpurun_out/bench_bst.txt 2>&1; echo "bst rc=$?" >> gpurun_out/bench_bst.txt
timeout 600 python tools/bench_cfg4.py > gpurun_out/bench_cfg4.txt 2>&1; echo "cfg4 rc=$?" >> gpurun_out/bench_cfg4.txt; cut -c1-260 gpurun_out/bench_cfg4.txt
cat gpurun_out/bench_bst.txt | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 6 -c 3 -f -o gpurun_out/tc_kernels python tools/run_xprop.py 0.25 3 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu.log
tail -2 gpurun_out/ncu.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.txt') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],'clocks',d['clocks'])
    print('roofline',{k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac','kernel')})
else:
    print(open('gpurun_out/bench.txt').read()[-3000:])
PY
