#!/bin/bash
# multi-GPU iteration: NCCL parity test, then bench.py at N GPUs: blocking all-reduce (round 1), side stream with SM margins
N=${1:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TEST" ]; then timeout 600 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q > gpurun_out/pytest_nccl.txt 2>&1; echo "nccl test rc=$?"; tail -1 gpurun_out/pytest_nccl.txt; fi
run() {  # name, env assignments...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 200 --warmup 5 --no-extras $BENCH_ARGS > gpurun_out/bench_${N}gpu_$name.txt 2> gpurun_out/bench_${N}gpu_$name.err
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/bench_${N}gpu_$name.txt") if x.startswith("{")]
    d=json.loads(l[-1]); print("$name", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), d["config"]["allreduce"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${N}gpu_$name.err").read()[-1500:])
PY
}
BENCH_ARGS="--blocking-allreduce" run blocking A=1
for m in ${MARGINS:-8 12}; do
  BENCH_ARGS="--sm-margin $m" run dyn_margin$m A=1
  BENCH_ARGS="--sm-margin $m" run static_margin$m BSMM_TILE_QUEUE=static
done
