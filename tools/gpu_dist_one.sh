#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_final.txt 2> gpurun_out/bench_${N}gpu_final.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_${N}gpu_final.err
python tools/show_bench.py gpurun_out/bench_${N}gpu_final.txt | sed -n 1,2p; python tools/show_bench.py gpurun_out/bench_${N}gpu_final.txt | tail -1
