"""Times fprop / bprop / updat at BASELINE cfg 2 over densities (rotating buffers, CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul, _lib
from bench import make_layout

import os
dens = [float(a) for a in sys.argv[1:]] or [0.05, 0.10, 0.25, 0.50, 1.00]
N, BS = 4096, int(os.environ.get("BS", "32"))
AXIS = int(os.environ.get("AXIS", "1"))
X = [(torch.randn((N, 4096), device="cuda") * 0.1).bfloat16() for _ in range(3)]
E = [(torch.randn((N, 4096), device="cuda") * 0.1).bfloat16() for _ in range(3)]


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for d in dens:
    bsmm = BlocksparseMatMul(make_layout(d, 4096 // BS, 4096 // BS), block_size=BS, feature_axis=AXIS)
    W = (torch.randn(bsmm.w_shape, device="cuda") * 0.01).bfloat16()
    fl = 2.0 * bsmm.blocks * BS * BS * N
    out = []
    for name, fn in [("fprop", lambda i: bsmm.fprop(X[i % 3], W)), ("bprop", lambda i: bsmm.bprop(E[i % 3], W)),
                     ("updat", lambda i: bsmm.updat([X[i % 3]], [E[i % 3]]))]:
        ms = timeit(fn)
        out.append("%s %.4f ms %6.1f TF (%s)" % (name, ms, fl / ms / 1e9, _lib.last_kernel()))
    print("axis %d bs %d density %.2f nnz %5d | " % (AXIS, BS, d, bsmm.blocks) + " | ".join(out), flush=True)
assert _lib.device_error() == 0
