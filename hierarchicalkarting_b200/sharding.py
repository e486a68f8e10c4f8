"""Multi-GPU sharding of the hot path: every LQNG problem and every rollout is independent, so each rank
(one process per GPU) takes a contiguous slice of the batch / a disjoint Philox counter range; there is NO collective on
the data path — only a final gather of per-rank summaries (a few numbers), done with torch.distributed (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` units for `rank`: sizes differ by at most one, earlier ranks take the extras."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rollout_shard(n_rollouts: int, rank: int, world: int) -> tuple[int, int]:
    """(rollout_offset, count) for hk_mcts_rollouts: disjoint Philox counter ranges, union = [0, n_rollouts)."""
    lo, hi = shard_range(n_rollouts, rank, world)
    return lo, hi - lo


def tree_shard(n_trees: int, seed: int, rank: int, world: int) -> tuple[int, int, int]:
    """(first tree, count, seed to pass) for hk_mcts_forest_search / hk_mcts_search_seq_batch on this rank: tree r of a call gets Philox key
    seed + r, so a rank that searches the global trees [lo, hi) passes seed + lo and every tree keeps the key it has in a one-rank call."""
    lo, hi = shard_range(n_trees, rank, world)
    return lo, hi - lo, (seed + lo) & 0xFFFFFFFFFFFFFFFF


def gather_sum(dist, summary, device=None):
    """Final gather of per-rank summaries: all_gather + sum (summaries are KB-sized; latency-bound, not bandwidth-bound)."""
    import torch
    t = torch.as_tensor(np.asarray(summary, dtype=np.float64), device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return t.cpu().numpy()
    parts = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.stack(parts).sum(dim=0).cpu().numpy()


def max_over_ranks(dist, value: float, device=None) -> float:
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
