"""Sharding of a parameter batch over the GPUs of one box (one process per GPU).

The hot path has no inter-evaluation dependency, so the batch is cut into contiguous row blocks,
one per rank, with NO collective on the data path; the only exchange is the final gather of the
per-vector results (nll, per-component nll, gradients) -- a few MiB at most.
``local_eval`` is whatever evaluates a block on this rank: in the product it is the CUDA handle
(``Handle.eval_batch``); the CPU (gloo) tests pass a stand-in.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(B: int, world_size: int):
    """Contiguous, balanced row blocks: the first ``B % world_size`` ranks get one extra row."""
    base, extra = divmod(int(B), int(world_size))
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def shard_range(B: int, rank: int, world_size: int):
    b = shard_bounds(B, world_size)
    return b[rank], b[rank + 1]


def eval_sharded(local_eval, P: np.ndarray, *, dist=None, device=None, gather=True):
    """Evaluate rows ``shard_range(B, rank, world)`` of ``P`` with ``local_eval(P_block) ->
    (nll[b], comp[b, C] | None, grad[b, K] | None)`` and, if ``gather``, all-gather the results so
    every rank returns full-batch arrays. ``dist`` is ``torch.distributed`` (already initialised)
    or None for a single process."""
    P = np.ascontiguousarray(np.atleast_2d(P), dtype=np.float64)
    B = P.shape[0]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_eval(P)
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(B, rank, world)
    nll, comp, grad = local_eval(P[lo:hi])
    if not gather:
        return nll, comp, grad
    bounds = shard_bounds(B, world)
    dev = device if device is not None else torch.device("cpu")

    def gather_rows(local, width):
        if local is None:
            return None
        local = np.ascontiguousarray(local, dtype=np.float64).reshape(hi - lo, width)
        # uneven blocks: pad to the largest block so all_gather sees equal shapes
        maxrows = max(bounds[r + 1] - bounds[r] for r in range(world))
        buf = torch.zeros((maxrows, width), dtype=torch.float64, device=dev)
        buf[: hi - lo] = torch.from_numpy(local).to(dev)
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        full = np.empty((B, width))
        for r in range(world):
            full[bounds[r]:bounds[r + 1]] = outs[r][: bounds[r + 1] - bounds[r]].cpu().numpy()
        return full

    full_nll = gather_rows(nll, 1)[:, 0]
    full_comp = gather_rows(comp, comp.shape[1]) if comp is not None else None
    full_grad = gather_rows(grad, grad.shape[1]) if grad is not None else None
    return full_nll, full_comp, full_grad
