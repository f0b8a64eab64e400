"""Property tests (hypothesis) for the host-side schedule builders: for ANY layout every block of the LUT is issued
exactly once, on the right accumulator, whatever the tile width / pipeline variant / CTA count."""
import numpy as np
from hypothesis import given, settings, strategies as st

from blocksparse_b200.lut import MatmulLuts, TransformerLuts, pick_tile_count
from oracle.bsmm_oracle import MatmulOracle


@st.composite
def layouts(draw, max_side=24):
    cb = draw(st.integers(1, max_side))
    kb = draw(st.integers(1, max_side))
    density = draw(st.sampled_from([0.05, 0.2, 0.5, 1.0]))
    seed = draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    lay = (rng.random((cb, kb)) < density).astype(np.int32)
    lay[rng.integers(cb), rng.integers(kb)] = 1          # the reference requires at least one block
    return lay


@settings(max_examples=40, deadline=None)
@given(layouts(), st.booleans())
def test_luts_equal_the_oracles(lay, z):
    L, O = MatmulLuts(lay, z_order=z), MatmulOracle(lay, 32, 1, z_order=z)
    np.testing.assert_array_equal(L.updat_lut, O.updat_lut)
    np.testing.assert_array_equal(L.fprop_lut, O.fprop_lut)
    np.testing.assert_array_equal(L.bprop_lut, O.bprop_lut)
    assert L.fprop_list == O.fprop_list and L.bprop_list == O.bprop_list


@settings(max_examples=40, deadline=None)
@given(layouts(), st.booleans(), st.sampled_from([(32, 8, 8), (32, 8, 4), (32, 8, 2), (32, 16, 8), (64, 8, 4), (64, 4, 2)]),
       st.integers(0, 3))
def test_tile_schedule_issues_every_block_once(lay, bprop, cfg, extra_tiles):
    bsize, T, wpg = cfg
    L = MatmulLuts(lay)
    lists = dict(L.bprop_list if bprop else L.fprop_list)
    n_out = L.CB if bprop else L.KB
    n_tiles = -(-n_out // T) + extra_tiles                # the host may ask for more, narrower tiles (pick_tile_count)
    if n_tiles > n_out:
        n_tiles = n_out
    s, off = L.tile_schedule(bprop, T, bsize, wpg, n_tiles=n_tiles)
    assert s[0] == n_tiles and off % 32 == 0
    wbytes16 = (bsize * bsize * 2) >> 4
    seen, next_out, next_group = set(), 0, 0
    for t in range(n_tiles):
        fg, ng, fo, packed = (int(v) for v in s[4 + 4 * t: 8 + 4 * t])
        no = packed & 0xff
        assert fg == next_group and fo == next_out and 0 <= no <= T
        next_out, next_group = fo + no, fg + ng
        for g in range(fg, fg + ng):
            rec = s[off + 32 * g: off + 32 * g + 32]
            ib, n_w, n_runs = int(rec[0]), int(rec[1]) & 0xff, int(rec[1]) >> 8
            assert 1 <= n_runs <= n_w <= wpg
            staged = 0
            for r in range(n_runs):
                r0, r1 = int(rec[12 + r]), int(rec[20 + r])
                w_slot, col, n = (r0 & 0xffff) // wbytes16, r0 >> 16, (r1 >> 17) << 3
                assert w_slot == staged and n % bsize == 0 and 0 < n <= 256 and (r1 & 1) == 1
                for i in range(n // bsize):
                    w, slot = int(rec[4 + w_slot + i]), col // bsize + i
                    assert slot < no and (ib, w) in lists[fo + slot] and w not in seen
                    seen.add(w)
                staged += n // bsize
            assert staged == n_w
    assert next_out == n_out and len(seen) == L.blocks


@settings(max_examples=40, deadline=None)
@given(layouts(max_side=40), st.sampled_from([32, 64]), st.sampled_from([None, 1, 7, 148]))
def test_updat_schedule_issues_every_block_once(lay, bsize, n_cta):
    L = MatmulLuts(lay)
    s, off = L.updat_schedule(bsize, n_cta=n_cta)
    n_tiles, G, KT = int(s[0]), int(s[1]), int(s[2])
    rec = s[off:].reshape(n_tiles, 64)
    assert (np.diff(rec[:, 1]) <= 0).all()
    seen, gk = set(), set()
    for t in range(n_tiles):
        c0, n_act = int(rec[t, 0]), int(rec[t, 1])
        assert c0 % G == 0 and 1 <= n_act <= KT
        ks = rec[t, 8:8 + n_act]
        assert (np.diff(ks) > 0).all()
        for sl, k in enumerate(ks.tolist()):
            assert (c0, k) not in gk
            gk.add((c0, k))
            ws = [int(rec[t, 16 + i * KT + sl]) for i in range(G) if rec[t, 16 + i * KT + sl] >= 0]
            assert ws
            for w in ws:
                c, kk = L.updat_lut[w]
                assert c0 <= c < c0 + G and kk == k and w not in seen
                seen.add(w)
    assert len(seen) == L.blocks


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 3), st.integers(1, 12), st.integers(0, 2 ** 16))
def test_nt_items_pair_every_block_once(heads, nb, seed):
    rng = np.random.default_rng(seed)
    lay = (rng.random((heads, nb, nb)) < 0.4).astype(np.int32)
    n = int(lay[0].sum())
    if n == 0:
        lay[:, 0, 0] = 1
        n = 1
    for h in range(1, heads):                               # the reference requires equal block counts per head
        lay[h] = 0
        idx = rng.permutation(nb * nb)[:n]
        lay[h].flat[idx] = 1
    L = TransformerLuts(lay, 64)
    for h in range(L.lut_heads):
        seen = set()
        for k_blk, nv, b0, q0, b1, q1, _, _ in L.nt_items[h].tolist():
            for b, q in [(b0, q0), (b1, q1)][:nv]:
                assert tuple(L.nt_lut[h][b]) == (q, k_blk) and b not in seen
                seen.add(b)
        assert len(seen) == L.blocks


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 512), st.integers(1, 64), st.sampled_from([148, 296]), st.sampled_from([4, 8, 16]))
def test_pick_tile_count_is_feasible(n_out, n_ntiles, slots, max_t):
    n_kt = pick_tile_count(n_out, n_ntiles, slots, max_t)
    assert -(-n_out // max_t) <= n_kt <= n_out              # tiles are at most max_t wide and never empty
