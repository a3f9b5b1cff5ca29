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


def tp_unique_id(dist=None, src: int = 0) -> bytes:
    """The 128-byte NCCL unique id of a tensor-parallel group: made on rank `src`
    (l3_nccl_unique_id) and broadcast over the caller's torch.distributed group (any backend).
    An id names exactly one communicator: call this once per `Llama` instance."""
    import ctypes as C
    from . import _cabi
    buf = C.create_string_buffer(128)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1 or dist.get_rank() == src:
        _cabi.check(_cabi.lib().l3_nccl_unique_id(buf))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return buf.raw
    box = [buf.raw if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def tp_shard_shapes(dim, n_heads, n_kv_heads, hidden_dim, vocab_size, world):
    """Per-rank shapes of the packed matrices under tensor parallelism: heads and
    FFN columns split for the column-parallel Wqkv / Wgate|Wup, input columns split for the
    row-parallel Wo / Wdown, vocabulary rows split for the LM head."""
    if n_kv_heads % world or hidden_dim % (8 * world) or vocab_size % world:
        raise ValueError(f"world {world} must divide n_kv_heads, hidden_dim/8 and vocab_size")
    hd = dim // n_heads
    hn, kv, fd = n_heads // world, n_kv_heads // world, hidden_dim // world
    return {"wqkv": ((hn + 2 * kv) * hd, dim), "wo": (dim, hn * hd), "w13": (2 * fd, dim),
            "w2": (dim, fd), "lm_head": (vocab_size // world, dim), "kv_cache_heads": kv}
