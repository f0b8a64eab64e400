"""Host-side logic that needs no GPU: updat schedules for every block size, tile order tables, synthetic layouts, the NCCL
SM-margin helper, argument validation of the utility functions."""
import os

import numpy as np
import pytest
import torch

from blocksparse_b200 import dist as bdist
from blocksparse_b200.layouts import barabasi_albert_layout, barabasi_albert_m, bernoulli_layout, local_strided_layout
from blocksparse_b200.lut import MatmulLuts, tile_order, updat_record_shape


@pytest.mark.parametrize("bs", [16, 32, 64])
@pytest.mark.parametrize("shape,density", [((9, 40), 0.3), ((33, 17), 0.15), ((12, 12), 1.0), ((64, 64), 0.05)])
def test_updat_schedule_covers_every_block_once(shape, density, bs):
    """build_updat_schedule: every weight block appears in exactly one tile, under the right (input block, slot), slots list the
    tile's kept output blocks, tiles are sorted longest first, the record layout matches csrc/tc_updat.cuh."""
    rng = np.random.default_rng(sum(shape) + bs)
    lay = (rng.random(shape) < density).astype(np.int32)
    lay[0, 0] = 1
    luts = MatmulLuts(lay)
    sched, off = luts.updat_schedule(bs, n_cta=148)
    n_tiles, G, KT, REC = (int(v) for v in sched[:4])
    assert (REC, 128 // bs, 256 // bs) == (updat_record_shape(bs)[0], G, KT)
    TAB = updat_record_shape(bs)[1]
    rec = sched[off:].reshape(n_tiles, REC)
    seen = np.zeros(luts.blocks, dtype=np.int64)
    n_act = rec[:, 1]
    assert np.all(np.diff(n_act) <= 0) and n_act.min() >= 1 and n_act.max() <= KT
    for r in rec:
        c0, na = int(r[0]), int(r[1])
        ks = r[8:8 + na]
        assert len(set(ks.tolist())) == na and np.all(ks >= 0)
        tab = r[TAB:TAB + G * KT].reshape(G, KT)
        assert np.all(tab[:, na:] == -1)
        for i in range(G):
            for s in range(na):
                w = int(tab[i, s])
                if w >= 0:
                    seen[w] += 1
                    assert tuple(luts.updat_lut[w]) == (c0 + i, int(ks[s]))
        assert np.all((tab[:, :na] >= 0).any(axis=0))          # every kept output block has at least one active block in the group
    assert np.all(seen == 1)


def test_tile_order_is_a_heaviest_first_permutation():
    cost = np.array([5.0, 40.0, 6.0, 5.5, 39.0])
    order = tile_order(cost, n_ntiles=3)
    assert sorted(order.tolist()) == list(range(15))
    kt = order % 5
    assert set(kt[:6].tolist()) == {1, 4}                      # the two heavy output tiles of all three minibatch tiles come first
    uniform = tile_order(np.full(7, 3.0), n_ntiles=4)
    assert uniform.tolist() == list(range(28))                  # equal costs keep the natural order


def test_synthetic_layouts():
    rng = np.random.default_rng(0)
    lay = bernoulli_layout(rng, 16, 24, 0.2)
    assert lay.shape == (16, 24) and all(lay[i, i] for i in range(16))
    for d in (0.1, 0.25):
        ba = barabasi_albert_layout(64, d, np.random.default_rng(1))
        m = barabasi_albert_m(64, d)
        assert ba.shape == (64, 64) and np.array_equal(ba, ba.T) and np.all(np.diag(ba) == 1) and np.all(ba[:m, :m] == 1)
        assert ba.sum() >= 0.9 * d * 64 * 64
        assert ba.sum(0).max() >= 2 * ba.sum(0).mean()          # skewed: a few block columns hold most of the blocks
    ls = local_strided_layout(64)
    assert ls.sum() == 453 and np.all(np.triu(ls, 1) == 0)


def test_reserve_sms_for_nccl_sets_the_environment(monkeypatch):
    for k in ("BSMM_SM_MARGIN", "NCCL_MAX_CTAS"):
        monkeypatch.delenv(k, raising=False)
    assert bdist.reserve_sms_for_nccl(12, nccl_ctas=8) == 12
    assert os.environ["BSMM_SM_MARGIN"] == "12" and os.environ["NCCL_MAX_CTAS"] == "8"
    assert bdist.reserve_sms_for_nccl(4) == 12                   # an explicit earlier setting wins
    monkeypatch.setenv("BSMM_SM_MARGIN", "0")
    monkeypatch.delenv("NCCL_MAX_CTAS", raising=False)
    assert bdist.reserve_sms_for_nccl(8) == 0 and "NCCL_MAX_CTAS" not in os.environ


def test_utility_argument_validation_without_a_gpu():
    from blocksparse_b200 import _lib, blocksparse_norm, blocksparse_prune, blocksparse_reduced_dw
    with pytest.raises(ValueError):
        blocksparse_norm(torch.zeros(3, 8, 16))
    with pytest.raises(_lib.BsmmError):
        blocksparse_norm(torch.zeros(3, 8, 8))                   # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        blocksparse_prune(torch.zeros(3, 8, 8), torch.ones(4), step=0, sparsity=0.5)
    with pytest.raises(ValueError):
        blocksparse_reduced_dw([torch.zeros(4, 4)] * 9, [torch.zeros(4, 4)] * 9, 1.0)
    with pytest.raises(_lib.BsmmError):
        blocksparse_reduced_dw([torch.zeros(32, 4).half()], [torch.zeros(32, 4).half()], 1.0)
