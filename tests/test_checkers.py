"""The op classes' NumPy checker methods (blocksparse_b200/checkers.py: the reference's fprop_test ... masked_softmax_grad_test)
against (a) the fixtures recorded from the reference itself and (b) the oracle's independent restatement.  No GPU needed."""
import os

import numpy as np
import pytest

from tests._util import GOLDEN, golden_files
from blocksparse_b200 import BlocksparseMatMul, BlocksparseTransformer
from oracle import wutil_oracle
from oracle.bsmm_oracle import MatmulOracle
from tests.golden.make_golden import causal_callback, checker_callback


@pytest.mark.parametrize("fname", golden_files("bsmm_"))
def test_matmul_checkers_reproduce_the_reference_outputs(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    bsmm = BlocksparseMatMul(g["layout"], block_size=int(g["bsize"]), feature_axis=int(g["axis"]))
    np.testing.assert_allclose(bsmm.fprop_test(g["X"], g["W"]), g["Y"], rtol=1e-6, atol=5e-6)
    np.testing.assert_allclose(bsmm.bprop_test(g["E"], g["W"]), g["DX"], rtol=1e-6, atol=5e-6)
    np.testing.assert_allclose(bsmm.updat_test(g["X"], g["E"]), g["DW"], rtol=1e-6, atol=5e-6)
    if "gate" in g.files:
        np.testing.assert_allclose(bsmm.fprop_test(g["X"], g["W"], gate=g["gate"]), g["Y_gated"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(bsmm.bprop_test(g["E"], g["W"], gate=g["gate"]), g["DX_gated"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(bsmm.updat_test(g["X"], g["E"], gate=g["gate"], dw_gated=True), g["DW_gated"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("fname", golden_files("wutil_"))
def test_l2_normalize_checkers_and_oracle_reproduce_the_reference(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    bs = int(g["bsize"])
    bsmm = BlocksparseMatMul(g["layout"], block_size=bs, feature_axis=0)
    orc = MatmulOracle(g["layout"], bs, 0)
    np.testing.assert_allclose(bsmm.l2_normalize_test(g["W"]), g["Y"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(bsmm.l2_normalize_grad_test(g["W"], g["U"]), g["DX"], rtol=2e-5, atol=2e-5)
    y, _ = wutil_oracle.l2_normalize(orc.fprop_list, g["W"], bs)
    dx, _ = wutil_oracle.l2_normalize_grad(orc.fprop_list, g["W"], g["U"], bs)
    np.testing.assert_allclose(y, g["Y"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(dx, g["DX"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("fname", golden_files("bst_"))
def test_transformer_checkers_reproduce_the_reference_outputs(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    cb = None
    if bool(g["has_mask"]):
        cb = causal_callback if "tril" in fname else checker_callback
    bst = BlocksparseTransformer(g["layout"], block_size=int(g["bs"]), heads=int(g["heads"]), mask_callback=cb)
    scale = float(g["scale"])
    S = bst.nt_test(g["Q"], g["K"])
    np.testing.assert_allclose(S, g["S"], rtol=1e-5, atol=1e-5)
    P = bst.masked_softmax_test(g["S"], scale=scale)
    np.testing.assert_allclose(P, g["P"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bst.nn_test(g["P"], g["V"]), g["Y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bst.tn_test(g["P"], g["DY"]), g["DV"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bst.masked_softmax_grad_test(g["DP"], g["P"], scale=scale), g["DS"], rtol=1e-4, atol=1e-5)
    if "P_auto" in g.files:
        Pa = bst.masked_softmax_test(g["S"], scale=scale, autoregress_at_key=int(g["autoregress_at_key"]))
        np.testing.assert_allclose(Pa, g["P_auto"], rtol=1e-5, atol=1e-6)
