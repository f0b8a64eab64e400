"""NumPy checker methods of the op classes (`fprop_test`, `nt_test`, ... in the reference: blocksparse/matmul.py:353-453,
blocksparse/transformer.py:186-305).

The reference carries these on the op objects as CPU test helpers; they are part of its public surface, so they exist
here too.  They are NOT a compute path: no op calls them, and they are written independently of `oracle/` (vectorised
einsum / segment sums over the LUT arrays instead of per-block loops), which makes `tests/test_checkers.py` a cross-check
of two restatements of the same reference code.
"""
import numpy as np


class MatmulCheckers(object):
    """Mixin for BlocksparseMatMul: needs updat_lut, bsize, axis, CB, KB, C, K, w_shape, fprop_list."""

    def _dense(self, W):
        bs = self.bsize
        D = np.zeros((self.CB, bs, self.KB, bs), dtype=np.float64)
        D[self.updat_lut[:, 0], :, self.updat_lut[:, 1], :] = W
        return D.reshape(self.C, self.K)

    def fprop_test(self, I, W, gate=None):
        """O = I . W over the active blocks (matmul.py:353-375)."""
        Wg = W if gate is None else W * np.asarray(gate, dtype=W.dtype).reshape(-1, 1, 1)
        D = self._dense(Wg)
        return I.astype(np.float64) @ D if self.axis else D.T @ I.astype(np.float64)

    def bprop_test(self, E, W, gate=None):
        """B = E . W^T (matmul.py:377-399)."""
        Wg = W if gate is None else W * np.asarray(gate, dtype=W.dtype).reshape(-1, 1, 1)
        D = self._dense(Wg)
        return E.astype(np.float64) @ D.T if self.axis else D @ E.astype(np.float64)

    def updat_test(self, I, E, gate=None, dw_gated=False):
        """U[w] = I[c-blk] . E[k-blk]^T over the minibatch (matmul.py:401-419)."""
        bs = self.bsize
        cs, ks = self.updat_lut[:, 0], self.updat_lut[:, 1]
        if self.axis:
            Iv = I.astype(np.float64).reshape(-1, self.CB, bs)
            Ev = E.astype(np.float64).reshape(-1, self.KB, bs)
            U = np.einsum('nbi,nbj->bij', Iv[:, cs, :], Ev[:, ks, :])
        else:
            Iv = I.astype(np.float64).reshape(self.CB, bs, -1)
            Ev = E.astype(np.float64).reshape(self.KB, bs, -1)
            U = np.einsum('bin,bjn->bij', Iv[cs], Ev[ks])
        if dw_gated and gate is not None:
            U = U * np.asarray(gate, dtype=np.float64).reshape(-1, 1, 1)
        return U

    def _column_ids(self):
        """block id -> output block column, and blocks sorted by column."""
        return self.updat_lut[:, 1].astype(np.int64)

    def l2_normalize_test(self, W, epsilon=1e-12):
        """matmul.py:421-429: every output feature of a block column is normalised over all rows of all its blocks."""
        col = self._column_ids()
        ss = np.zeros((self.KB, self.bsize), dtype=np.float64)
        np.add.at(ss, col, np.square(W.astype(np.float64)).sum(axis=1))
        norm = np.sqrt(np.maximum(ss, epsilon))
        return (W / norm[col][:, None, :]).astype(W.dtype)

    def l2_normalize_grad_test(self, W, U, epsilon=1e-12):
        """matmul.py:431-443."""
        col = self._column_ids()
        W64, U64 = W.astype(np.float64), U.astype(np.float64)
        ss = np.zeros((self.KB, self.bsize), dtype=np.float64)
        np.add.at(ss, col, np.square(W64).sum(axis=1))
        mx = np.maximum(ss, epsilon)
        red = np.zeros_like(ss)
        np.add.at(red, col, (-U64 * W64).sum(axis=1))
        red = red / mx * (ss >= epsilon)
        return ((U64 + W64 * red[col][:, None, :]) / np.sqrt(mx)[col][:, None, :]).astype(U.dtype)


class TransformerCheckers(object):
    """Mixin for BlocksparseTransformer: needs nt_lut, heads, lut_heads, blk_size, blocks, ctx_blks_q/k, softmax_mask_np."""

    def _head_lut(self, h):
        return self.nt_lut[h if self.lut_heads > 1 else 0]

    def _split_heads(self, X, ctx_blks):
        B, _, S = X.shape
        return X.reshape(B, ctx_blks, self.blk_size, self.heads, S // self.heads)

    def nt_test(self, A, B):
        """C[n,h,b] = A[n,q-blk,:,h,:] . B[n,k-blk,:,h,:]^T (transformer.py:186-203)."""
        Av, Bv = self._split_heads(A, self.ctx_blks_q), self._split_heads(B, self.ctx_blks_k)
        C = np.empty((A.shape[0], self.heads, self.blocks, self.blk_size, self.blk_size), dtype=np.float32)
        for h in range(self.heads):
            lut = self._head_lut(h)
            C[:, h] = np.einsum('nbid,nbjd->nbij', Av[:, :, :, h, :][:, lut[:, 0]], Bv[:, :, :, h, :][:, lut[:, 1]])
        return C

    def _xn_check(self, A, B, out_col, in_col, n_out, transpose):
        Bv = self._split_heads(B, self.ctx_blks_q if transpose else self.ctx_blks_k)
        nb, S = B.shape[0], B.shape[2]
        C = np.zeros((nb, n_out, self.blk_size, self.heads, S // self.heads), dtype=np.float32)
        for h in range(self.heads):
            lut = self._head_lut(h)
            Ah = A[:, h].astype(np.float32)
            if transpose:
                Ah = Ah.transpose(0, 1, 3, 2)
            prod = np.einsum('nbij,nbjd->nbid', Ah, Bv[:, :, :, h, :][:, lut[:, in_col]])
            for n in range(nb):
                Ch = np.zeros((n_out,) + prod.shape[2:], dtype=np.float32)
                np.add.at(Ch, lut[:, out_col], prod[n])
                C[n, :, :, h, :] = Ch
        return C.reshape(nb, n_out * self.blk_size, S)

    def nn_test(self, A, B):
        """C[n,q-blk] += A[n,h,b] . B[n,k-blk] (transformer.py:205-223)."""
        return self._xn_check(A, B, 0, 1, self.ctx_blks_q, False)

    def tn_test(self, A, B):
        """C[n,k-blk] += A[n,h,b]^T . B[n,q-blk] (transformer.py:225-243)."""
        return self._xn_check(A, B, 1, 0, self.ctx_blks_k, True)

    def _visible(self, h, autoregress_at_key=None):
        """bool [blocks, bs, bs]: key j of block b visible to query r (bit j of mask word r; transformer.py:262-279)."""
        bs = self.blk_size
        if self.softmax_mask_np is None:
            return np.ones((self.blocks, bs, bs), dtype=bool)
        hl = h if self.lut_heads > 1 else 0
        words = self.softmax_mask_np[hl].astype(np.uint64)                    # [blocks, bs]
        if autoregress_at_key is not None:
            lut = self._head_lut(h).astype(np.int64)
            q0, k0 = lut[:, 0] * bs, lut[:, 1] * bs
            r = np.arange(bs, dtype=np.int64)
            sa = bs - np.clip(autoregress_at_key - k0, 0, bs)                 # [blocks]
            sb = np.clip(bs - 1 + k0[:, None] - (q0[:, None] + r[None, :]), 0, bs)
            shift = np.minimum(sa[:, None], sb).astype(np.uint64)
            ones = np.uint64((1 << bs) - 1) if bs < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
            shifted = np.where(shift >= 64, np.uint64(0), ones >> np.minimum(shift, np.uint64(63)))
            words = words & shifted
        j = np.arange(bs, dtype=np.uint64)
        return ((words[:, :, None] >> j[None, None, :]) & np.uint64(1)).astype(bool)

    def masked_softmax_test(self, x, scale=1.0, autoregress_at_key=None):
        """Row softmax over all the blocks of a query-block row; masked entries count as -FLT_MAX (transformer.py:246-286)."""
        y = np.empty_like(x)
        neg = -np.finfo(np.float32).max
        bs = self.blk_size
        for h in range(self.heads):
            q = self._head_lut(h)[:, 0]
            vis = self._visible(h, autoregress_at_key)
            xm = np.where(vis[None], x[:, h].astype(np.float32) * np.float32(scale), np.float32(neg))
            mx = np.full((x.shape[0], self.ctx_blks_q, bs), neg, dtype=np.float32)
            for n in range(x.shape[0]):
                np.maximum.at(mx[n], q, xm[n].max(axis=2))
            e = np.exp(xm - mx[:, q][..., None])
            sm = np.zeros((x.shape[0], self.ctx_blks_q, bs), dtype=np.float32)
            for n in range(x.shape[0]):
                np.add.at(sm[n], q, e[n].sum(axis=2))
            y[:, h] = e / sm[:, q][..., None]
        return y

    def masked_softmax_grad_test(self, dy, y, scale=1.0):
        """dx = (dy - sum_row(dy * y)) * y * scale (transformer.py:289-305)."""
        dx = np.empty_like(dy)
        bs = self.blk_size
        for h in range(self.heads):
            q = self._head_lut(h)[:, 0]
            prod = (dy[:, h] * y[:, h]).sum(axis=3)
            tot = np.zeros((dy.shape[0], self.ctx_blks_q, bs), dtype=prod.dtype)
            for n in range(dy.shape[0]):
                np.add.at(tot[n], q, prod[n])
            dx[:, h] = (dy[:, h] - tot[:, q][..., None]) * y[:, h] * scale
        return dx
