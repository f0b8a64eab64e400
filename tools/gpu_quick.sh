#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/bench1.txt 2>&1; echo "rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench1.txt') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']); print('cpu',d['cpu_baseline'])
else:
    print(open('gpurun_out/bench1.txt').read()[-3000:])
PY
cat /sys/bus/pci/devices/*/local_cpulist 2>/dev/null | sort | uniq -c | head -5; nvidia-smi topo -m 2>/dev/null | head -8
