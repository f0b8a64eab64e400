#!/bin/bash
# Quick perf iteration: TC parity tests, bench sweep, one ncu capture of the xprop kernel.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/pytest_tc.txt 2>&1; rc=$?; echo "pytest_tc rc=$rc" >> gpurun_out/pytest_tc.txt
tail -5 gpurun_out/pytest_tc.txt
if [ $rc -ne 0 ]; then echo "TC tests failed: stopping early"; tail -30 gpurun_out/pytest_tc.txt; exit 1; fi
timeout 600 python bench.py --steps 10 --warmup 3 --sweep --no-cpu > gpurun_out/bench.txt 2>&1; echo "bench rc=$?" >> gpurun_out/bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_xprop -s 2 -c 1 -f -o gpurun_out/xprop python tools/run_xprop.py 0.25 2 > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu.log
tail -2 gpurun_out/ncu.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.txt') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',d['value'],'ms/step',d['ms_per_step'],'per_op',d['roofline']['per_op_ms'])
    for k,v in d.get('density_sweep',{}).items():
        print(k, {o:(round(v[o]['ms'],4), round(v[o]['tflops'],1)) for o in ('fprop','bprop','updat')})
else:
    print(open('gpurun_out/bench.txt').read()[-2000:])
PY
