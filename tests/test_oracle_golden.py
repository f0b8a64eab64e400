"""Pins the CPU oracle (oracle/*.py) against fixtures produced by the reference itself.

The fixtures come from tests/golden/make_golden.py, which imports the reference's
blocksparse/matmul.py and transformer.py (TensorFlow mocked) and records their LUTs
and NumPy `*_test` outputs.  Integer artefacts must match bit-exactly; float outputs
to float32 round-off (both sides are NumPy on the same inputs).
"""
import os

import numpy as np
import pytest

from tests._util import GOLDEN, golden_files
from oracle.bsmm_oracle import MatmulOracle, z_order_2d, fprop_fast, bprop_fast, updat_fast
from oracle.bst_oracle import TransformerOracle
from tests.golden.make_golden import causal_callback, checker_callback


def test_z_order_known_values():
    # utils.py:95-103 -- x on even bits, y on odd bits
    assert z_order_2d(0, 0) == 0
    assert z_order_2d(1, 0) == 1
    assert z_order_2d(0, 1) == 2
    assert z_order_2d(3, 5) == 0b100111
    assert z_order_2d(127, 127) == (1 << 14) - 1


@pytest.mark.parametrize("fname", golden_files("bsmm_"))
def test_matmul_oracle_matches_reference(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    orc = MatmulOracle(g["layout"], int(g["bsize"]), int(g["axis"]))
    np.testing.assert_array_equal(orc.fprop_lut, g["fprop_lut"])
    np.testing.assert_array_equal(orc.bprop_lut, g["bprop_lut"])
    np.testing.assert_array_equal(orc.updat_lut, g["updat_lut"])
    meta = [orc.fprop_segments, orc.fprop_locks, orc.fprop_shared,
            orc.bprop_segments, orc.bprop_locks, orc.bprop_shared, orc.blocks, orc.C, orc.K]
    np.testing.assert_array_equal(np.array(meta), g["meta"])
    # the as-is (row-major scipy.find) reference run must agree on block ids and math
    np.testing.assert_array_equal(orc.updat_lut, g["asis_updat_lut"])

    W, X, E = g["W"], g["X"], g["E"]
    np.testing.assert_allclose(orc.fprop(X, W), g["Y"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(orc.bprop(E, W), g["DX"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(orc.updat(X, E), g["DW"], rtol=1e-5, atol=2e-6)
    # dense einsum restatement agrees with the block loops
    np.testing.assert_allclose(orc.fprop_dense(X, W), g["Y"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(orc.bprop_dense(E, W), g["DX"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(orc.updat_dense(X, E), g["DW"], rtol=1e-4, atol=1e-5)
    # the BLAS-batched variant timed as cpu_baseline computes the same thing (fp32)
    np.testing.assert_allclose(fprop_fast(orc, X, W), g["Y"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(bprop_fast(orc, E, W), g["DX"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(updat_fast(orc, X, E), g["DW"], rtol=2e-4, atol=2e-4)
    if "gate" in g.files:
        gate = g["gate"]
        np.testing.assert_allclose(orc.fprop(X, W, gate=gate), g["Y_gated"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(orc.bprop(E, W, gate=gate), g["DX_gated"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(orc.updat(X, E, gate=gate, dw_gated=True), g["DW_gated"], rtol=1e-5, atol=2e-6)


def _callback_for(name, has_mask):
    if not has_mask:
        return None
    return checker_callback if "perhead" in name else causal_callback


@pytest.mark.parametrize("fname", golden_files("bst_"))
def test_transformer_oracle_matches_reference(fname):
    g = np.load(os.path.join(GOLDEN, fname))
    cb = _callback_for(fname, bool(g["has_mask"]))
    orc = TransformerOracle(g["layout"], int(g["bs"]), heads=int(g["heads"]), mask_callback=cb)
    np.testing.assert_array_equal(orc.nt_lut, g["nt_lut"])
    np.testing.assert_array_equal(orc.nn_lut, g["nn_lut"])
    np.testing.assert_array_equal(orc.tn_lut, g["tn_lut"])
    np.testing.assert_array_equal(
        np.array([orc.blocks, orc.nn_max, orc.tn_max, orc.ctx_blks_q, orc.ctx_blks_k]), g["meta"])
    if cb is not None:
        np.testing.assert_array_equal(orc.softmax_mask_np, g["mask_np"])
        np.testing.assert_array_equal(orc.softmax_mask, g["mask_dev"])
    Q, K, V, DY = g["Q"], g["K"], g["V"], g["DY"]
    scale = float(g["scale"])
    S = orc.nt(Q, K)
    np.testing.assert_allclose(S, g["S"], rtol=1e-6, atol=1e-6)
    P = orc.masked_softmax(g["S"], scale=scale)
    np.testing.assert_allclose(P, g["P"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.nn(g["P"], V), g["Y"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.tn(g["P"], DY), g["DV"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.nt(DY, V), g["DP"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.masked_softmax_grad(g["DP"], g["P"], scale=scale), g["DS"], rtol=1e-5, atol=1e-6)
    if cb is not None:
        ak = int(g["autoregress_at_key"])
        np.testing.assert_allclose(orc.masked_softmax(g["S"], scale=scale, autoregress_at_key=ak),
                                   g["P_auto"], rtol=1e-6, atol=1e-7)
    # dense attention cross-check of the whole NT -> softmax -> NN chain
    Yd = orc.dense_attention(Q, K, V, scale=scale)
    np.testing.assert_allclose(g["Y"], Yd, rtol=1e-4, atol=1e-5)
