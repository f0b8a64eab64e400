"""Host (CPU) cost of enqueueing the bench step: perf_counter around 300 steps of fprop + bprop + updat without any sync,
next to the GPU time of the same steps, and the same step replayed from a CUDA graph."""
import sys, time
import torch
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul, _lib
from bench import make_layout
for dens in (0.05, 0.25):
    bsmm = BlocksparseMatMul(make_layout(dens), block_size=32, feature_axis=1)
    N = 4096
    W = (torch.randn(bsmm.w_shape, device="cuda") * 0.01).bfloat16()
    X = (torch.randn(bsmm.i_shape(N), device="cuda") * 0.1).bfloat16()
    E = (torch.randn(bsmm.o_shape(N), device="cuda") * 0.1).bfloat16()
    def step():
        bsmm.fprop(X, W); bsmm.bprop(E, W); bsmm.updat([X], [E])
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 300
    a.record(); t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize()
    c, d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c.record(); t2 = time.perf_counter()
    for _ in range(n):
        g.replay()
    t3 = time.perf_counter(); d.record(); torch.cuda.synchronize()
    print("density %.2f: eager  CPU enqueue %.4f ms/step (%.1f us per launch), GPU %.4f ms/step | CUDA graph replay CPU %.4f ms/step, GPU %.4f ms/step"
          % (dens, (t1 - t0) / n * 1e3, (t1 - t0) / n / 3 * 1e6, a.elapsed_time(b) / n, (t3 - t2) / n * 1e3, c.elapsed_time(d) / n))
assert _lib.device_error() == 0
