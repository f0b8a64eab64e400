"""Pipeline trace of the pair-schedule xprop kernel (BSMM_TRACE=1): per-group clock64 deltas of CTA 0."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul, _lib
from bench import make_layout
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
bsmm = BlocksparseMatMul(make_layout(d), block_size=32, feature_axis=1)
X = (torch.randn((4096, 4096), device="cuda") * 0.1).bfloat16()
W = (torch.randn(bsmm.w_shape, device="cuda") * 0.01).bfloat16()
for _ in range(3):
    y = bsmm.fprop(X, W)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 2048)()
_lib.check(_lib.load().bsmm_debug_trace(buf, 2048), "trace")
t = np.array(buf, dtype=np.int64).reshape(256, 8)[:, :5]
t0 = t[t > 0].min()
print("kernel", _lib.last_kernel(), "density", d)
print("grp   P_free  P_issued  I_full  I_turn  I_done   | full-issued  turn-full  done-turn  next P_free - I_done(g-NP)")
for g in range(64):
    r = t[g] - t0
    print("%3d  %7d %7d %7d %7d %7d   | %6d %6d %6d" % (g, r[0], r[1], r[2], r[3], r[4], r[2] - r[1], r[3] - r[2], r[4] - r[3]))
