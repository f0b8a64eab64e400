"""world_size-2 gloo test of the N>1 host logic (no GPU): shard the minibatch, all-reduce the partial dW.

The per-rank math is done by the oracle here (the product has no CPU path); what is under test is the sharding /
reduction wiring in blocksparse_b200/dist.py: sum over ranks of updat(shard) == updat(full minibatch), and
fprop/bprop of a shard == the corresponding slice of the full result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blocksparse_b200 import dist as bdist
from oracle.bsmm_oracle import MatmulOracle


def test_shard_bounds_partition_the_minibatch():
    for N in (1, 7, 64, 4097):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_bounds(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, axis, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    lay = (rng.random((5, 6)) < 0.5).astype(np.int32)
    lay[0, 0] = 1
    orc = MatmulOracle(lay, 8 if axis == 0 else 32, axis)
    N = 37                                           # not divisible by the world size
    W = rng.normal(0, 0.1, orc.w_shape).astype(np.float32)
    X = rng.normal(0, 1, orc.i_shape(N)).astype(np.float32)
    E = rng.normal(0, 1, orc.o_shape(N)).astype(np.float32)
    xs = bdist.shard_minibatch(torch.as_tensor(X), axis).numpy()
    es = bdist.shard_minibatch(torch.as_tensor(E), axis).numpy()
    # fprop / bprop are independent per column: the shard result is a slice of the full result
    a, b = bdist.shard_bounds(N, rank, world)
    y_full, dx_full = orc.fprop(X, W), orc.bprop(E, W)
    ys, dxs = orc.fprop(xs, W), orc.bprop(es, W)
    sl = (slice(None), slice(a, b)) if axis == 0 else (slice(a, b), slice(None))
    ok = np.allclose(ys, y_full[sl], atol=1e-5) and np.allclose(dxs, dx_full[sl], atol=1e-5)
    # updat: partial dW per rank, all-reduced
    dw = torch.as_tensor(orc.updat(xs, es).astype(np.float32))
    bdist.allreduce_dw(dw)
    ok = ok and np.allclose(dw.numpy(), orc.updat(X, E), atol=1e-4)
    avg = torch.as_tensor(orc.updat(xs, es).astype(np.float32))
    bdist.allreduce_dw(avg, average=True)
    ok = ok and np.allclose(avg.numpy() * world, orc.updat(X, E), atol=1e-4)
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("axis", [0, 1])
def test_sharded_updat_allreduce_matches_full_batch(axis):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, axis, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
