"""Batch-sharded multi-GPU driver.

Frames are independent (``z_run`` starts at zero every call,
Evaluation/dac_vcpwq_proposed6_latency.py:461), so the batch is split into contiguous
shards, one process per GPU, weights replicated, and NO collective runs on the hot
path.  The only exchange is the final gather of the code indices (a few KB per frame)
and metric scalars.  The reference itself is single-process, single-device.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous, balanced split: the first (n_items % world) ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_ragged(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """all_gather of per-rank tensors whose dim 0 differs (counts[r] rows on rank r);
    returns the concatenation in rank order on every rank."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


class ShardedCodec:
    """Runs ``forward_eval`` on this rank's shard of a global batch and gathers the indices.

    ``forward_fn(a, t, books_use) -> (y, idx)`` is the per-GPU hot path (ProposedEval on CUDA in
    production; any callable in the CPU tests)."""

    def __init__(self, forward_fn, group=None):
        self.forward_fn = forward_fn
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def local_slice(self, n_items: int):
        return shard_bounds(n_items, self.world, self.rank)

    def run_local(self, a_local: torch.Tensor, t_local: torch.Tensor, books_use=None):
        """The hot path on this rank's own shard: no communication at all.  -> (y_local, idx_local)."""
        return self.forward_fn(a_local, t_local, books_use)

    def gather_indices(self, idx_local: torch.Tensor, counts=None) -> torch.Tensor:
        """The path's only exchange: every rank's code indices, concatenated in rank order (called once, after the
        last step).  counts[r] = rows of rank r (default: equal shards)."""
        if self.world == 1:
            return idx_local
        if counts is None:
            counts = [idx_local.shape[0]] * self.world
        return gather_ragged(idx_local, counts, self.group)

    def run(self, a_global: torch.Tensor, t_global: torch.Tensor, books_use=None, gather_y: bool = False):
        """a_global/t_global: the full [B, 1, T] batch, identical on every rank (or only this rank's
        slice is read).  Returns (y_local, idx_global[, y_global])."""
        b = a_global.shape[0]
        lo, hi = self.local_slice(b)
        y, idx = self.forward_fn(a_global[lo:hi], t_global[lo:hi], books_use)
        if self.world == 1:
            return (y, idx, y) if gather_y else (y, idx)
        counts = [shard_bounds(b, self.world, r)[1] - shard_bounds(b, self.world, r)[0] for r in range(self.world)]
        idx_all = gather_ragged(idx, counts, self.group)
        if gather_y:
            return y, idx_all, gather_ragged(y, counts, self.group)
        return y, idx_all


def codec_forward_fn(net):
    """Adapter: ProposedEval -> forward_fn for ShardedCodec."""

    def fn(a, t, books_use):
        y = net.forward_eval(a, t, books_use)
        return y, net.last_indices

    return fn
