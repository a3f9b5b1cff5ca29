"""Weight ingestion: Hugging Face `safetensors` Llama checkpoints -> the reference's
`.npz` key layout (`/root/reference/llama3.py:219-235, 269, 280-281`: HF-style names, float32, `[out, in]`).

The key names already coincide.  What differs is the ROW ORDER of q_proj / k_proj: HF checkpoints are
permuted for the half-split (`rotate_half`) RoPE, the reference rotates INTERLEAVED pairs
(`apply_rotary_emb`, llama3.py:41-76).  `unpermute_qk=True` (default) undoes the HF permutation so that
the reference arithmetic - and therefore this library - reproduces the checkpoint's function.  Tied
embeddings (`lm_head.weight` absent) are materialised.  The reference ignores `rope_theta` (base 10000);
Llama-3 checkpoints use 500000: pass `honor_rope_theta=True` to `Llama` to opt in.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, Mapping, Union

import numpy as np


def hf_permute(w: np.ndarray, n_heads: int) -> np.ndarray:
    """Interleaved-pair row order -> HF half-split order (what HF's conversion script applies)."""
    d1, d2 = w.shape
    return w.reshape(n_heads, d1 // n_heads // 2, 2, d2).transpose(0, 2, 1, 3).reshape(d1, d2)


def hf_unpermute(w: np.ndarray, n_heads: int) -> np.ndarray:
    """HF half-split row order -> interleaved pairs (the layout the reference's RoPE expects)."""
    d1, d2 = w.shape
    return w.reshape(n_heads, 2, d1 // n_heads // 2, d2).transpose(0, 2, 1, 3).reshape(d1, d2)


def _to_f32(a) -> np.ndarray:
    if hasattr(a, "detach"):  # torch tensor (bf16 shards)
        import torch
        return a.detach().to(torch.float32).cpu().numpy()
    return np.asarray(a, dtype=np.float32)


def convert_state_dict(tensors: Mapping[str, object], n_heads: int, n_kv_heads: int, unpermute_qk: bool = True) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    for key, val in tensors.items():
        if key.endswith("rotary_emb.inv_freq"):
            continue
        w = _to_f32(val)
        if unpermute_qk and re.search(r"self_attn\.q_proj\.weight$", key):
            w = hf_unpermute(w, n_heads)
        elif unpermute_qk and re.search(r"self_attn\.k_proj\.weight$", key):
            w = hf_unpermute(w, n_kv_heads)
        out[key] = np.ascontiguousarray(w, dtype=np.float32)
    if "lm_head.weight" not in out and "model.embed_tokens.weight" in out:
        out["lm_head.weight"] = out["model.embed_tokens.weight"]  # tie_word_embeddings
    return out


def safetensors_to_npz(paths: Union[str, Iterable[str]], dst_npz: str, n_heads: int, n_kv_heads: int,
                       unpermute_qk: bool = True) -> Dict[str, tuple]:
    """Convert one or more safetensors shards into one `.npz` in the reference layout; returns the shapes."""
    from safetensors import safe_open
    if isinstance(paths, str):
        paths = [paths]
    tensors = {}
    for p in paths:
        try:
            with safe_open(p, framework="np") as f:
                for k in f.keys():
                    tensors[k] = f.get_tensor(k)
        except Exception:  # bf16 shards are not representable in NumPy: go through torch
            with safe_open(p, framework="pt") as f:
                for k in f.keys():
                    tensors[k] = f.get_tensor(k)
    out = convert_state_dict(tensors, n_heads, n_kv_heads, unpermute_qk)
    np.savez(dst_npz, **out)
    return {k: v.shape for k, v in out.items()}
