"""GPU parity: BlocksparseTransformer through the C ABI vs the oracle / reference fixtures.

fp32 inputs: NT/NN/TN run true-fp32 FMA (<=1e-5 when the sparse operand is fp32).  The public
query_key_op emits bf16 scores exactly like the reference (transformer.py:343), so chains through it
are held to the 16-bit tolerance 1e-2.
"""
import os

import numpy as np
import pytest
import torch

from tests._util import GOLDEN, golden_files, ref_errors
from tests.golden.make_golden import causal_callback, checker_callback
from blocksparse_b200 import BlocksparseTransformer
from oracle.bst_oracle import TransformerOracle

pytestmark = pytest.mark.gpu


def cb_for(name, has_mask):
    if not has_mask:
        return None
    return checker_callback if "perhead" in name else causal_callback


def close(got, ref, tol, what, abs_tol=None):
    """tol = (max|d|/mean|ref|, ||d||2/||ref||2), the reference's two metrics.  For probabilities and scores the
    mean is tiny compared with the largest element (softmax rows are peaked), so the max metric is replaced by an
    absolute bound `abs_tol` on values that live in [0, 1] / O(1)."""
    g = got.detach().float().cpu().numpy()
    mx, l2 = ref_errors(g, ref)
    if abs_tol is not None:
        worst = float(np.abs(g - ref).max())
        assert worst <= abs_tol and l2 <= tol[1], "%s: max_abs_err %.3e l2_err %.3e" % (what, worst, l2)
    else:
        assert mx <= tol[0] and l2 <= tol[1], "%s: max_err %.3e l2_err %.3e" % (what, mx, l2)


def rounded(a, dtype):
    t = torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dtype)
    return t.cuda(), t.float().numpy()


@pytest.mark.parametrize("fname", golden_files("bst_"))
def test_golden_fp32_ops(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    bst = BlocksparseTransformer(g["layout"], int(g["bs"]), heads=int(g["heads"]), mask_callback=cb_for(fname, bool(g["has_mask"])))
    scale = float(g["scale"])
    Q, K, V, DY = (torch.as_tensor(g[k]).cuda() for k in ("Q", "K", "V", "DY"))
    tol = (5e-5, 1e-5)       # fp32 FMA, different summation order than NumPy
    S = bst._nt(Q, K, torch.float32)
    close(S, g["S"], tol, "nt")
    P = bst._softmax(torch.as_tensor(g["S"]).cuda(), scale, bst.softmax_mask_np is not None, None, torch.float32)
    close(P, g["P"], (2e-5, 1e-5), "softmax")
    Pg = torch.as_tensor(g["P"]).cuda()
    close(bst._xn(Pg, V, False), g["Y"], tol, "nn")
    close(bst._xn(Pg, DY, True), g["DV"], tol, "tn")
    close(bst._nt(DY, V, torch.float32), g["DP"], tol, "nt(dy,v)")
    close(bst._softmax_grad(torch.as_tensor(g["DP"]).cuda(), Pg, scale), g["DS"], (2e-5, 1e-5), "softmax grad")
    if bool(g["has_mask"]):
        ak = int(g["autoregress_at_key"])
        Pa = bst._softmax(torch.as_tensor(g["S"]).cuda(), scale, True, ak, torch.float32)
        close(Pa, g["P_auto"], (2e-5, 1e-5), "softmax autoregress")
        # the standalone mask rewrite gives the same thing as the fused path
        orc = TransformerOracle(g["layout"], int(g["bs"]), heads=int(g["heads"]), mask_callback=cb_for(fname, True))
        bs = int(g["bs"])
        new_mask = bst.partial_autoregressive_mask(ak).cpu().numpy().view(orc.softmax_mask_np.dtype).reshape(orc.softmax_mask_np.shape)
        for hl in range(orc.lut_heads):
            for b, (q, k) in enumerate(orc.nt_list[hl]):
                vis = orc._mask_bits(hl, b, k, ak)
                words = (vis.astype(np.uint64) << np.arange(bs, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
                np.testing.assert_array_equal(new_mask[hl, b].astype(np.uint64), words)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("fname", golden_files("bst_"))
def test_public_chain_with_autograd(fname, dtype):
    """q,k,v -> query_key_op -> masked_softmax -> weight_value_op, forward and backward."""
    g = np.load(os.path.join(GOLDEN, fname))
    cb = cb_for(fname, bool(g["has_mask"]))
    bst = BlocksparseTransformer(g["layout"], int(g["bs"]), heads=int(g["heads"]), mask_callback=cb)
    orc = TransformerOracle(g["layout"], int(g["bs"]), heads=int(g["heads"]), mask_callback=cb)
    scale = float(g["scale"])
    (Qd, Qh), (Kd, Kh), (Vd, Vh), (Ed, Eh) = (rounded(g[k], dtype) for k in ("Q", "K", "V", "DY"))
    Qd.requires_grad_(); Kd.requires_grad_(); Vd.requires_grad_()
    w = bst.query_key_op(Qd, Kd)
    assert w.dtype == torch.bfloat16
    p = bst.masked_softmax(w, scale=scale)
    assert p.dtype == (torch.bfloat16 if dtype == torch.float32 else dtype)
    y = bst.weight_value_op(p, Vd)
    y.backward(Ed)
    # oracle chain in fp32 on the rounded inputs
    S = orc.nt(Qh, Kh)
    P = orc.masked_softmax(S, scale=scale)
    Y = orc.nn(P, Vh)
    DV = orc.tn(P, Eh)
    DP = orc.nt(Eh, Vh)
    DS = orc.masked_softmax_grad(DP, P, scale=scale)
    DQ = orc.nn(DS, Kh)
    DK = orc.tn(DS, Qh)
    tol = (4e-2, 1e-2)
    close(w, S, tol, "scores", abs_tol=2.0 ** -8 * float(np.abs(S).max()))      # bf16 scores: half an ulp of the largest
    close(p, P, tol, "probs", abs_tol=2.0 ** -7)                                 # values in [0,1], 16-bit storage
    # dense outputs: the l2 metric is the north-star tolerance; the max metric (worst element over MEAN magnitude)
    # carries the 2^-9 rounding of the 16-bit probabilities / scores times max/mean of the data
    close(y, Y, (1.5e-1, 1e-2), "y")
    close(Vd.grad, DV, (1.5e-1, 1e-2), "dv")
    close(Qd.grad, DQ, (2e-1, 2e-2), "dq")       # three 16-bit roundings deep
    close(Kd.grad, DK, (2e-1, 2e-2), "dk")


def test_cfg3_shape_properties():
    """BASELINE cfg 3 layout (ctx 4096, bs 64, local+strided causal): softmax rows sum to 1,
    masked keys get exactly 0, and the sparse chain equals dense causal-masked attention on a sample."""
    nb = 64
    q, k = np.indices((nb, nb))
    lay = ((k <= q) & ((q - k < 4) | (k % 8 == 7))).astype(np.int32)
    assert lay.sum() == 453
    heads, hs, bs, batch = 4, 64, 64, 1
    bst = BlocksparseTransformer(lay, bs, heads=heads, mask_callback=causal_callback)
    assert (bst.blocks, bst.nn_max, bst.tn_max) == (453, 11, 57)
    gen = torch.Generator(device="cuda").manual_seed(0)
    Q, K, V = ((torch.rand((batch, nb * bs, heads * hs), generator=gen, device="cuda") * 2 - 1).half() for _ in range(3))
    scale = 1.0 / np.sqrt(hs)
    w = bst.query_key_op(Q, K)
    p = bst.masked_softmax(w, scale=scale)
    y = bst.weight_value_op(p, V)
    # rows sum to one
    pf = p.float()
    row_sum = torch.zeros(batch, heads, nb, bs, device="cuda")
    qs = torch.as_tensor(bst.nt_lut[0][:, 0].astype(np.int64)).cuda()
    row_sum.index_add_(2, qs, pf.sum(-1))
    assert float((row_sum - 1).abs().max()) < 2e-2
    # diagonal blocks are causal: strictly-upper entries are exactly zero
    diag = torch.as_tensor(np.nonzero(bst.nt_lut[0][:, 0] == bst.nt_lut[0][:, 1])[0]).cuda()
    upper = torch.triu(torch.ones(bs, bs, device="cuda"), 1).bool()
    assert float(pf[:, :, diag][..., upper].abs().max()) == 0.0
    # dense reference on head 0 / first 1024 queries
    n = 1024
    Qh, Kh, Vh = (t[0, :, :hs].float() for t in (Q, K, V))
    s = (Qh[:n] @ Kh.t()) * scale
    vis = torch.as_tensor(np.kron(lay, np.ones((bs, bs), np.int32)))[:n].cuda().bool()
    vis &= torch.tril(torch.ones(n, nb * bs, device="cuda")).bool()
    s = s.masked_fill(~vis, float("-inf"))
    ref = torch.softmax(s, -1) @ Vh
    got = y[0, :n, :hs].float()
    assert float((got - ref).norm() / ref.norm()) < 2e-2
