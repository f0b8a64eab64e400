#!/bin/bash
# what the driver runs at N GPUs: reference arm, then bench.py with default flags; plus the 2-rank NCCL parity test
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -q 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_${N}gpu_ref.txt 2>&1; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_${N}gpu_ref.txt | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu_default.txt 2> gpurun_out/bench_${N}gpu_default.err; echo "bench rc=$?"; tail -c 500 gpurun_out/bench_${N}gpu_default.err
python tools/show_bench.py gpurun_out/bench_${N}gpu_default.txt
