"""blocksparse_b200 -- B200-native block-sparse matmul / block-sparse attention ops.

Drop-in for the hot path of openai/blocksparse: `BlocksparseMatMul` (fprop / bprop /
updat, group_param_grads) and `BlocksparseTransformer` (NT / NN / TN + masked softmax),
implemented as hand-written sm_100a CUDA behind the C ABI in include/bsmm_b200.h.
"""
from .matmul import (BlocksparseMatMul, SparseProj, block_reduced_full_dw, blocksparse_reduced_dw, group_param_grads)
from .optimize import blocksparse_l2_decay, blocksparse_norm, blocksparse_prune
from .transformer import BlocksparseTransformer
from .lut import z_order_2d
from . import _lib

__version__ = "0.1.0"
__all__ = ["BlocksparseMatMul", "BlocksparseTransformer", "SparseProj", "group_param_grads", "blocksparse_reduced_dw",
           "block_reduced_full_dw", "blocksparse_norm", "blocksparse_prune", "blocksparse_l2_decay", "z_order_2d"]
