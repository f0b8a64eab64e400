#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $N --steps 100 --warmup 5 --no-extras > gpurun_out/bench_${N}gpu_rep$i.txt 2> gpurun_out/bench_${N}gpu_rep$i.err
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/bench_${N}gpu_rep$i.txt") if x.startswith("{")]
    d=json.loads(l[-1]); print("rep$i", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), d["config"]["allreduce"][:60])
except Exception as e:
    print("rep$i FAILED", e); print(open("gpurun_out/bench_${N}gpu_rep$i.err").read()[-1200:])
PY
done
