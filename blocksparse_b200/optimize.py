"""Block-sparse weight utilities of the reference's blocksparse/optimize.py (tail, :294-335) on torch tensors.

  blocksparse_norm(param, norm="max")                       per-block max|w| or l2 norm -> float32 [blocks]
  blocksparse_l2_decay(param, gate=None, rate, epsilon)     in place: w -= w * min(rate / sqrt(sum w^2 + eps), 1)
  blocksparse_prune(param, gate, step, sparsity= | threshold=, norm, frequency)
                                                            in place on `gate`; top-k by block norm or threshold
All run as hand-written CUDA kernels (csrc/wutil.cuh) through the C ABI; there is no CPU path.
"""
import numpy as np
import torch

from . import _lib


def _check_param_shape(param, gate=None):
    if param.dim() != 3 or param.shape[1] != param.shape[2] or param.shape[1] not in (8, 16, 32, 64):
        raise ValueError("param must be (blocks, bsize, bsize) with bsize in {8,16,32,64}, got %s" % (tuple(param.shape),))
    if gate is not None and (gate.dim() != 1 or gate.shape[0] != param.shape[0] or gate.dtype != torch.float32):
        raise ValueError("gate must be a float32 vector with one entry per block")
    if not param.is_cuda or not param.is_contiguous():
        raise _lib.BsmmError("block-sparse utilities need contiguous CUDA tensors (no CPU path)")


def blocksparse_norm(param, norm="max"):
    _check_param_shape(param)
    out = torch.empty(param.shape[0], dtype=torch.float32, device=param.device)
    with torch.cuda.device(param.device):
        rc = _lib.load().bsmm_block_norm(_lib.dtype_code(param.dtype), param.shape[1], param.shape[0], param.data_ptr(),
                                         out.data_ptr(), 1 if norm.lower() == "l2" else 0, _lib.stream_ptr())
    _lib.check(rc, "bsmm_block_norm")
    return out


def blocksparse_l2_decay(param, gate=None, rate=0.05, epsilon=1e-12):
    """In place, like the reference's op (it aliases the variable); returns `param`."""
    _check_param_shape(param, gate)
    with torch.cuda.device(param.device):
        rc = _lib.load().bsmm_l2_decay(_lib.dtype_code(param.dtype), param.shape[1], param.shape[0], param.data_ptr(),
                                       _lib.ptr(gate), float(rate), float(epsilon), _lib.stream_ptr())
    _lib.check(rc, "bsmm_l2_decay")
    return param


def blocksparse_prune(param, gate, step, sparsity=None, threshold=None, norm="max", frequency=1):
    """Update `gate` in place every `frequency` steps: keep the (1 - sparsity) share of blocks with the largest norm, or
    the blocks whose norm reaches `threshold` (optimize.py:319-339).  Returns `gate`."""
    _check_param_shape(param, gate)
    assert (sparsity is None) ^ (threshold is None), "exactly one of sparsity / threshold must be set"
    if int(step) % int(frequency) != 0:
        return gate
    lib = _lib.load()
    blocks = param.shape[0]
    with torch.cuda.device(param.device):
        if sparsity is not None:
            keep_frac = np.float32(1.0) - np.float32(sparsity)
            if keep_frac > 1.0:                       # negative sparsity makes this a no-op (optimize_op.cc:659-660)
                return gate
            norms = blocksparse_norm(param, norm=norm)
            idx = torch.argsort(norms, descending=True, stable=True).to(torch.int32)
            keep = int(np.float32(blocks) * keep_frac + np.float32(0.5))        # optimize_op.cc:663
            rc = lib.bsmm_prune_topk(gate.data_ptr(), idx.data_ptr(), blocks, keep, _lib.stream_ptr())
            _lib.check(rc, "bsmm_prune_topk")
        else:
            rc = lib.bsmm_threshold_prune(_lib.dtype_code(param.dtype), param.shape[1], blocks, param.data_ptr(), gate.data_ptr(),
                                          float(threshold), 1 if norm.lower() == "l2" else 0, _lib.stream_ptr())
            _lib.check(rc, "bsmm_threshold_prune")
    return gate
