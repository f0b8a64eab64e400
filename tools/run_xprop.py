"""Runs a few tcgen05 xprop launches of BASELINE cfg 2 (for ncu captures)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul
from bench import make_layout
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bsmm = BlocksparseMatMul(make_layout(d), block_size=32, feature_axis=1)
W = (torch.randn(bsmm.w_shape, device="cuda") * 0.01).bfloat16()
X = (torch.randn(bsmm.i_shape(4096), device="cuda") * 0.1).bfloat16()
E = (torch.randn(bsmm.o_shape(4096), device="cuda") * 0.1).bfloat16()
for _ in range(reps):
    bsmm.fprop(X, W)
    bsmm.bprop(E, W)
    bsmm.updat([X], [E])
torch.cuda.synchronize()
