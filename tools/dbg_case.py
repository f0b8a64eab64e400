import sys, numpy as np, torch
sys.path.insert(0, ".")
from blocksparse_b200 import BlocksparseMatMul, _lib
CB, KB, density, N, bs = [(5, 37, 0.5, 200, 32), (8, 8, 0.3, 128, 32), (64, 64, 0.2, 640, 32)][int(sys.argv[1]) if len(sys.argv) > 1 else 0]
rng = np.random.default_rng(1)
lay = rng.random((CB, KB)) < density
lay[0, 0] = True
b = BlocksparseMatMul(lay, block_size=bs, feature_axis=1)
W = torch.as_tensor(rng.normal(0, 0.1, b.w_shape).astype(np.float32)).bfloat16().cuda()
X = torch.as_tensor(rng.normal(0, 1, b.i_shape(N)).astype(np.float32)).bfloat16().cuda()
y = b.fprop(X, W, flags=_lib.FLAG_FORCE_TC)
print("device_error", _lib.device_error(), _lib.load().bsmm_last_error())
yr = b.fprop(X, W, flags=_lib.FLAG_FORCE_GENERIC)
print("maxdiff", (y.float() - yr.float()).abs().max().item())
E = torch.as_tensor(rng.normal(0, 1, b.o_shape(N)).astype(np.float32)).bfloat16().cuda()
dx = b.bprop(E, W, flags=_lib.FLAG_FORCE_TC)
print("bprop device_error", _lib.device_error(), _lib.load().bsmm_last_error())
dxr = b.bprop(E, W, flags=_lib.FLAG_FORCE_GENERIC)
print("maxdiff", (dx.float() - dxr.float()).abs().max().item())
