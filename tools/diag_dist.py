"""torchrun diagnostic: where does the multi-GPU step lose time?  GPU ms/step and CPU enqueue ms/step for the bench
step with (A) no all-reduce, (B) blocking all-reduce, (C) side-stream all-reduce, (D) async_op all-reduce."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul
from blocksparse_b200 import dist as bdist
from bench import make_layout

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl")
dev = torch.device("cuda", lr)
bsmm = BlocksparseMatMul(make_layout(0.25), block_size=32, feature_axis=1)
N = 4096
W = (torch.randn(bsmm.w_shape, device=dev) * 0.01).bfloat16()
Xs = [(torch.randn(bsmm.i_shape(N), device=dev) * 0.1).bfloat16() for _ in range(3)]
Es = [(torch.randn(bsmm.o_shape(N), device=dev) * 0.1).bfloat16() for _ in range(3)]
red = bdist.AllreduceStream(dev)
works = []


def step(i, mode):
    x, e = Xs[i % 3], Es[i % 3]
    bsmm.fprop(x, W); bsmm.bprop(e, W)
    dw = bsmm.updat([x], [e])
    if mode == "B":
        dist.all_reduce(dw)
    elif mode == "C":
        red.reduce(dw)
        if len(red.pending) >= 4:
            red.wait()
    elif mode == "D":
        works.append((dist.all_reduce(dw, async_op=True), dw))
        if len(works) >= 4:
            for w, _ in works:
                w.wait()
            works.clear()


def drain(mode):
    if mode == "C":
        red.wait()
    if mode == "D":
        for w, _ in works:
            w.wait()
        works.clear()


for mode in "ABCDA":
    for i in range(10):
        step(i, mode)
    drain(mode)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    for i in range(200):
        step(i, mode)
    drain(mode)
    cpu = (time.perf_counter() - t0) / 200 * 1e3
    b.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("mode %s: gpu %.4f ms/step, cpu enqueue %.4f ms/step" % (mode, a.elapsed_time(b) / 200, cpu), flush=True)
dist.destroy_process_group()
