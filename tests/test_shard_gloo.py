"""N > 1 host logic on CPU: two gloo ranks shard a batch, evaluate their block and all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gadget3_b200.shard import eval_sharded, shard_bounds, shard_range


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 8, 65536, 1000003):
        for w in (1, 2, 3, 8):
            b = shard_bounds(B, w)
            assert b[0] == 0 and b[-1] == B and len(b) == w + 1
            sizes = np.diff(b)
            assert sizes.max() - sizes.min() <= 1
            assert all(shard_range(B, r, w) == (b[r], b[r + 1]) for r in range(w))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gadget3_b200.configs import CONFIGS, parameter_batch
        from oracle.oracle_lib import Oracle
        model, p = CONFIGS["C1"]()
        orc = Oracle(*model.pack(p))
        full = np.tile(p.values(), (B, 1))
        full[:, p.optimised_indices()] = parameter_batch(p, B, seed=21)
        tidx = p.optimised_indices()[:3].astype(np.int32)
        seen = []

        def local_eval(block):            # stand-in for Handle.eval_batch on this rank's GPU
            seen.append(block.shape[0])
            return orc.eval_batch(block, tangent_idx=tidx, want_comp=True)

        nll, comp, grad = eval_sharded(local_eval, full, dist=dist)
        q.put((rank, seen[0], nll, comp, grad))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather(oracle_lib):
    from gadget3_b200.configs import CONFIGS, parameter_batch
    from oracle.oracle_lib import Oracle
    B, world = 7, 2                      # uneven split: 4 + 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    model, p = CONFIGS["C1"]()
    orc = Oracle(*model.pack(p))
    full = np.tile(p.values(), (B, 1))
    full[:, p.optimised_indices()] = parameter_batch(p, B, seed=21)
    tidx = p.optimised_indices()[:3].astype(np.int32)
    nll, comp, grad = orc.eval_batch(full, tangent_idx=tidx, want_comp=True)
    assert sorted(g[1] for g in got) == [3, 4]
    for rank, _, n2, c2, g2 in got:      # every rank holds the full, identical result
        assert (n2 == nll).all() and (c2 == comp).all() and (g2 == grad).all()
