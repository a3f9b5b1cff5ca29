"""Data-parallel plumbing for independent prompts (stories15M configs): one
process per GPU, full weights on each, prompts split in contiguous row blocks, NO data-path
collective.  torch.distributed is used only to synchronise and to combine timings / results."""
from __future__ import annotations

import numpy as np


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_rows` prompts owned by `rank`; blocks differ by <= 1 row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_prompts(ids: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_rows(ids.shape[0], rank, world)
    return np.ascontiguousarray(ids[lo:hi])


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Max of a per-rank scalar (timed region: the slowest rank defines the step)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_tokens(tokens: np.ndarray, dist=None):
    """All ranks' [rows_r, n] token blocks concatenated in rank order (rank 0 gets the result)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tokens
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, tokens)
    return np.concatenate(out, axis=0)
