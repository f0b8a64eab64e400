"""blocksparse_b200 -- B200-native block-sparse matmul / block-sparse attention ops.

Drop-in for the hot path of openai/blocksparse: `BlocksparseMatMul` (fprop / bprop /
updat, group_param_grads) and `BlocksparseTransformer` (NT / NN / TN + masked softmax),
implemented as hand-written sm_100a CUDA behind the C ABI in include/bsmm_b200.h.
"""
from .matmul import BlocksparseMatMul, group_param_grads
from .transformer import BlocksparseTransformer
from .lut import z_order_2d
from . import _lib

__version__ = "0.1.0"
__all__ = ["BlocksparseMatMul", "BlocksparseTransformer", "group_param_grads", "z_order_2d"]
