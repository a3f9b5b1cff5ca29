"""Random-init weights in the reference's `.npz` key layout.

The stories15M checkpoint is absent from the reference checkout
(`/root/reference/.MISSING_LARGE_BLOBS:1`) and no Llama-3 checkpoint is reachable, so every
parity test and benchmark runs on seeded synthetic weights written with exactly the keys
and `[out, in]` float32 shapes the reference reads (`llama3.py:219-235, 269, 280-281`).
Scales follow SURVEY.md §8(d): embeddings sigma 0.5, linears ~0.85/sqrt(fan_in), norm
weights 1 + 0.1*N(0,1) - logits stay O(1) and top-1/top-2 gaps stay well above fp32 noise.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Tuple

import numpy as np

from .config import ModelArgs


def weight_shapes(args: ModelArgs, hidden_dim: int) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (key, shape, kind) for every tensor of the layout; kind in embed|linear|norm."""
    d = args.dim
    hn = args.n_heads
    kvhn = hn if args.n_kv_heads is None else args.n_kv_heads
    hd = d // hn
    yield "model.embed_tokens.weight", (args.vocab_size, d), "embed"
    for i in range(args.n_layers):
        p = f"model.layers.{i}."
        yield p + "self_attn.q_proj.weight", (hn * hd, d), "linear"
        yield p + "self_attn.k_proj.weight", (kvhn * hd, d), "linear"
        yield p + "self_attn.v_proj.weight", (kvhn * hd, d), "linear"
        yield p + "self_attn.o_proj.weight", (d, hn * hd), "linear"
        yield p + "mlp.up_proj.weight", (hidden_dim, d), "linear"
        yield p + "mlp.gate_proj.weight", (hidden_dim, d), "linear"
        yield p + "mlp.down_proj.weight", (d, hidden_dim), "linear"
        yield p + "input_layernorm.weight", (d,), "norm"
        yield p + "post_attention_layernorm.weight", (d,), "norm"
    yield "model.norm.weight", (d,), "norm"
    yield "lm_head.weight", (args.vocab_size, d), "linear"


def make_weights(args: ModelArgs, hidden_dim: int, seed: int = 0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for key, shape, kind in weight_shapes(args, hidden_dim):
        if kind == "embed":
            w = 0.5 * rng.standard_normal(shape, dtype=np.float32)
        elif kind == "linear":
            w = (0.85 / math.sqrt(shape[1])) * rng.standard_normal(shape, dtype=np.float32)
        else:
            w = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
        out[key] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def save_npz(path: str, weights: Dict[str, np.ndarray]) -> None:
    """Write the mapping uncompressed, as `np.load` (reference `utils.py:4-5`) expects."""
    np.savez(path, **weights)


def param_count(args: ModelArgs, hidden_dim: int) -> int:
    return sum(int(np.prod(s)) for _, s, _ in weight_shapes(args, hidden_dim))
