"""GPU parity of the weight-format utilities (csrc/wutil.cuh) against the oracle restatements and the reference fixtures:
l2_normalize (+ gain, + gradients), block norm / l2 decay / pruning, identity_init, block-reduced full dW, SparseProj."""
import os

import numpy as np
import pytest
import torch

from tests._util import GOLDEN, golden_files, ref_errors
from blocksparse_b200 import (BlocksparseMatMul, SparseProj, block_reduced_full_dw, blocksparse_l2_decay, blocksparse_norm,
                              blocksparse_prune, blocksparse_reduced_dw)
from oracle import wutil_oracle
from oracle.bsmm_oracle import MatmulOracle

pytestmark = pytest.mark.gpu


def close(got, ref, tol, what):
    mx, l2 = ref_errors(got.detach().float().cpu().numpy() if torch.is_tensor(got) else got, ref)
    assert l2 <= tol, "%s: l2 %.3e max %.3e" % (what, l2, mx)


@pytest.mark.parametrize("fname", golden_files("wutil_"))
def test_l2_normalize_matches_reference_fixture(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    bs = int(g["bsize"])
    bsmm = BlocksparseMatMul(g["layout"], block_size=bs, feature_axis=0)
    W = torch.as_tensor(g["W"]).cuda().requires_grad_()
    y = bsmm.l2_normalize(W)
    close(y, g["Y"], 1e-6, "l2_normalize")
    y.backward(torch.as_tensor(g["U"]).cuda())
    close(W.grad, g["DX"], 2e-6, "l2_normalize grad")
    # columns of the sparse matrix have unit norm
    col = torch.as_tensor(bsmm.updat_lut[:, 1].astype(np.int64)).cuda()
    ss = torch.zeros(bsmm.KB, bs, device="cuda").index_add_(0, col, (y.detach() ** 2).sum(1))
    live = torch.as_tensor(np.asarray(g["layout"]).sum(0) > 0).cuda()
    assert float((ss[live] - 1).abs().max()) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("bs", [8, 32, 64])
def test_l2_normalize_with_gain_and_gradients(bs, dtype):
    rng = np.random.default_rng(bs)
    lay = (rng.random((5, 6)) < 0.5).astype(np.int32); lay[0, 0] = 1
    bsmm = BlocksparseMatMul(lay, block_size=bs, feature_axis=1)
    orc = MatmulOracle(lay, 32, 1)
    Wt = torch.as_tensor(rng.normal(0, 1, bsmm.w_shape).astype(np.float32)).to(dtype)
    Ut = torch.as_tensor(rng.normal(0, 1, bsmm.w_shape).astype(np.float32)).to(dtype)
    gain = rng.uniform(0.5, 2.0, bsmm.K).astype(np.float32)
    Wn, Un = Wt.float().numpy(), Ut.float().numpy()
    W = Wt.cuda().requires_grad_()
    G = torch.as_tensor(gain).cuda().requires_grad_()
    y = bsmm.l2_normalize(W, gain=G, dtype=torch.float32)
    yref, _ = wutil_oracle.l2_normalize(orc.fprop_list, Wn, bs, gain=gain)
    close(y, yref, 1e-6, "y")
    y.backward(Ut.cuda().float())
    dx, dg = wutil_oracle.l2_normalize_grad(orc.fprop_list, Wn, Un, bs, gain=gain)
    close(W.grad, dx, 1e-5 if dtype == torch.float32 else 4e-3, "dx")
    live = np.repeat(lay.sum(0) > 0, bs)
    close(G.grad[torch.as_tensor(live).cuda()], dg[:bsmm.K][live], 1e-5, "dg")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bs", [8, 16, 32, 64])
def test_block_norm_decay_and_pruning(bs, dtype):
    rng = np.random.default_rng(10 + bs)
    blocks = 37
    Wt = torch.as_tensor(rng.normal(0, 1, (blocks, bs, bs)).astype(np.float32) * rng.uniform(0.01, 1, (blocks, 1, 1)).astype(np.float32)).to(dtype)
    Wn = Wt.float().numpy()
    W = Wt.cuda()
    for norm in ("max", "l2"):
        close(blocksparse_norm(W, norm=norm), wutil_oracle.block_norm(Wn, norm), 1e-6, "norm " + norm)
        thr = float(np.median(wutil_oracle.block_norm(Wn, norm)))
        gate = torch.ones(blocks, device="cuda")
        blocksparse_prune(W, gate, step=4, threshold=thr * 1.0001, norm=norm, frequency=2)
        assert np.array_equal(gate.cpu().numpy(), wutil_oracle.threshold_prune(Wn, thr * 1.0001, norm))
        gate = torch.ones(blocks, device="cuda")
        blocksparse_prune(W, gate, step=3, threshold=thr, norm=norm, frequency=2)         # not a pruning step
        assert float(gate.min()) == 1.0
        gate = torch.ones(blocks, device="cuda")
        blocksparse_prune(W, gate, step=0, sparsity=0.3, norm=norm)
        assert np.array_equal(gate.cpu().numpy(), wutil_oracle.prune_topk(wutil_oracle.block_norm(Wn, norm), 0.3))
        assert int(gate.sum()) == int(np.float32(blocks) * np.float32(0.7) + np.float32(0.5))
    gate_np = (rng.random(blocks) < 0.7).astype(np.float32)
    Wd = Wt.clone().cuda()
    blocksparse_l2_decay(Wd, gate=torch.as_tensor(gate_np).cuda(), rate=0.05)
    close(Wd, wutil_oracle.l2_decay(Wn, gate_np, 0.05), {torch.float32: 1e-6, torch.float16: 1e-3, torch.bfloat16: 6e-3}[dtype], "l2_decay")
    assert torch.equal(Wd[torch.as_tensor(gate_np == 0).cuda()], W[torch.as_tensor(gate_np == 0).cuda()])   # gated blocks untouched


@pytest.mark.parametrize("shape", [(5, 5), (4, 7), (9, 3)])
def test_identity_and_ortho_init(shape):
    rng = np.random.default_rng(sum(shape))
    lay = (rng.random(shape) < 0.6).astype(np.int32)
    for i in range(min(shape)):
        lay[i, i] = 1
    bsmm = BlocksparseMatMul(lay, block_size=16, feature_axis=0)
    W = bsmm.identity_init(scale=0.5, dtype=torch.bfloat16)
    ref = wutil_oracle.identity_init(bsmm.updat_list, bsmm.CB, bsmm.KB, 16, 0.5)
    assert np.array_equal(W.float().cpu().numpy(), ref)
    gate = bsmm.checker_init()
    assert np.array_equal(gate.cpu().numpy(), np.array([(c & 1) ^ (k & 1) ^ 1 for c, k in bsmm.updat_list], dtype=np.float32))
    Wo = bsmm.ortho_init(rng=np.random.default_rng(0)).float().cpu().numpy()
    for k, col in bsmm.fprop_list:          # columns inside every block column are orthonormal when the column is tall enough
        if len(col) >= 1:
            M = np.concatenate([Wo[w] for _, w in col], axis=0)
            np.testing.assert_allclose(M.T @ M, np.eye(16), atol=1e-4)
    # prune(): reference semantics -- (new_param, new_gate), layout updated in place
    g2 = gate.clone(); g2[0] = 0
    new_w, new_g = bsmm.prune(W, g2)
    assert new_w.shape[0] == int((g2 != 0).sum()) == new_g.shape[0] and float(new_g.min()) == 1.0
    c, k = bsmm.updat_list[0]
    assert not bsmm.layout[c, k]


@pytest.mark.parametrize("axis,bs", [(0, 8), (0, 32), (1, 32), (1, 64)])
@pytest.mark.parametrize("norm", ["max", "l2"])
def test_block_reduced_full_dw(axis, bs, norm):
    rng = np.random.default_rng(bs + axis)
    bx, by, N, depth = 6, 4, 72, 11
    shape_x = (bx * bs, N) if axis == 0 else (N, bx * bs)
    shape_y = (by * bs, N) if axis == 0 else (N, by * bs)
    XS = [rng.normal(0, 1, shape_x).astype(np.float16).astype(np.float32) for _ in range(depth)]
    YS = [rng.normal(0, 1, shape_y).astype(np.float16).astype(np.float32) for _ in range(depth)]
    scale = 1.0 / (N * depth)
    xs = [torch.as_tensor(x).half().cuda() for x in XS]
    ys = [torch.as_tensor(y).half().cuda() for y in YS]
    dw, xr, yr = blocksparse_reduced_dw(xs[:8], ys[:8], scale, bsize=bs, norm=norm, axis=axis)
    DW, XR, YR = wutil_oracle.reduced_dw(XS[:8], YS[:8], scale, bs, axis, norm)
    close(xr, XR, 1e-3, "x_red"); close(yr, YR, 1e-3, "y_red"); close(dw, DW, 2e-3, "dw")
    full = block_reduced_full_dw(list(zip(xs, ys)), scale=scale, norm=norm, group_size=8, bsize=bs, axis=axis)   # 8 + 3 pairs, accumulated
    DWf, _, _ = wutil_oracle.reduced_dw(XS, YS, scale, bs, axis, norm)
    close(full, DWf, 2e-3, "grouped dw")


def test_sparse_proj_gather_scatter_and_gradients():
    rng = np.random.default_rng(3)
    nh, N = 96, 40
    sp = SparseProj(nh, proj_stride=3, block_size=8)
    assert sp.nproj == 32 and np.array_equal(sp.gather_lut, np.arange(0, 96, 3))
    x = torch.as_tensor(rng.normal(0, 1, (nh, N)).astype(np.float32)).cuda().requires_grad_()
    y = torch.as_tensor(rng.normal(0, 1, (sp.nproj, N)).astype(np.float32)).cuda().requires_grad_()
    gl = torch.as_tensor(sp.gather_lut.astype(np.int64)).cuda()
    g = sp.gather(x)
    assert torch.equal(g, x.detach()[gl])
    s = sp.scatter(y)
    ref = torch.zeros(nh, N, device="cuda"); ref[gl] = y.detach()
    assert torch.equal(s, ref)
    (g.sum() * 2 + (s * s).sum()).backward()
    gx = torch.zeros(nh, N, device="cuda"); gx[gl] = 2.0
    assert torch.equal(x.grad, gx) and torch.allclose(y.grad, 2 * y.detach())
    x.grad = y.grad = None
    za = sp.scatter_add(x, y)
    ra = x.detach().clone(); ra[gl] += y.detach()
    assert torch.equal(za, ra)
    zm = sp.scatter_mul(x, y)
    rm = x.detach().clone(); rm[gl] *= y.detach()
    assert torch.equal(zm, rm)
    (za.sum() + (zm * 3).sum()).backward()
    ex = torch.ones(nh, N, device="cuda") + 3; ex[gl] = 1 + 3 * y.detach()
    assert torch.allclose(x.grad, ex) and torch.allclose(y.grad, 1 + 3 * x.detach()[gl])
