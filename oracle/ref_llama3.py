"""ORACLE - test infrastructure, not product code.

A CPU (NumPy) restatement of the algorithm of the reference's Llama-3 forward pass and
greedy generate loop (`/root/reference/llama3.py:22-321`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module; the product path (`llama3.np_b200/`) never does and has no CPU fallback.

Pinning: this restatement is checked bit-for-bit against the *unmodified* reference,
imported from `/root/reference` in the authoring container, by `oracle/gen_golden.py`;
that script also writes the golden fixtures under `tests/golden/` that the CPU test-suite
replays (the reference itself holds no absolute golden vectors - its tests are
implementation-vs-implementation only, SURVEY.md section 4 - and its stories15M checkpoint
is absent, so fixtures are reference outputs on seeded random weights).

Behaviours of the reference reproduced on purpose:
  * float64 activations: the KV caches (`np.zeros`, llama3.py:138-153) and RoPE tables
    (llama3.py:31-38) are float64, weights stay float32; everything downstream of the
    first RoPE / first residual add is float64, logits are returned as float64.
  * `generate` feeds decode step i>=1 with `pos = L + i` (llama3.py:312-318): cache slot L
    is never written, stays zero and is attended.
  * `max_new_tokens` bounds the TOTAL length: `range(L, max_new_tokens)`.
  * RoPE base is 10000 regardless of `ModelArgs.rope_theta` (llama3.py:31, 272-274).
"""
from __future__ import annotations

import math
from typing import Mapping, Optional

import numpy as np


# ----------------------------------------------------------------------------- elementwise
def softmax_lastdim(x: np.ndarray) -> np.ndarray:
    """Row softmax with max subtraction - llama3.py:22-24."""
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def silu(x: np.ndarray) -> np.ndarray:
    """x * sigmoid(x) written as the reference does - llama3.py:27-28."""
    return x * (1 / (1 + np.exp(-x)))


def rms_norm(x: np.ndarray, weight: np.ndarray, eps: float) -> np.ndarray:
    """llama3.py:111-114: x / sqrt(mean(x^2) + eps) * w, dtype follows x."""
    ms = (x ** 2).mean(-1, keepdims=True) + eps
    return (x / np.sqrt(ms)) * weight


# ----------------------------------------------------------------------------------- RoPE
def rope_tables(head_dim: int, max_seq_len: int, base: int = 10000):
    """float64 cos/sin tables [M, HD/2] - llama3.py:31-38."""
    exponents = np.arange(0, head_dim, 2)[: head_dim // 2] / head_dim
    inv_freq = 1.0 / (base ** exponents)
    angles = np.outer(np.arange(max_seq_len), inv_freq)
    return np.cos(angles), np.sin(angles)


def rotate_pairs(x: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """Interleaved-pair rotation of one tensor [B, L, H, HD] - llama3.py:41-76.

    Element 2j is the 'real' and 2j+1 the 'imaginary' part, rotated by the angle of
    (position, j).  The reference rotates q and k in one call; one tensor at a time is
    the same arithmetic.
    """
    pairs = x.reshape(x.shape[:-1] + (-1, 2))
    re = pairs[..., 0]
    im = pairs[..., 1]
    c = cos[None, :, None, :]
    s = sin[None, :, None, :]
    out_re = re * c - im * s
    out_im = re * s + im * c
    return np.stack([out_re, out_im], axis=-1).reshape(out_re.shape[:-1] + (-1,))


# ---------------------------------------------------------------------------------- model
class OracleLlama:
    """Same public surface as the reference `Llama` (llama3.py:264-321)."""

    def __init__(self, weights, args, precast: bool = False):
        """`precast=True` (large test shapes only) converts to float64 ONCE every weight that the reference
        multiplies with float64 activations - NumPy would otherwise re-cast it on every call (the 83 % of
        SURVEY.md section 0).  The copy is C-contiguous like the temporary NumPy makes itself, so BLAS sees the same operands and the results are bit-identical
        (tests/test_oracle_cpu.py); layer 0's q / k / v projections keep float32 weights: their input
        is still float32 there (llama3.py:166-168 before the first RoPE)."""
        if isinstance(weights, str):
            weights = np.load(weights)  # utils.py:4-5
        self.args = args
        self.n_heads = args.n_heads
        self.n_kv_heads = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
        assert self.n_heads % self.n_kv_heads == 0  # llama3.py:127
        self.n_rep = self.n_heads // self.n_kv_heads
        self.head_dim = args.dim // args.n_heads
        self.embed = weights["model.embed_tokens.weight"]
        self.cos, self.sin = rope_tables(self.head_dim, args.max_seq_len)  # llama3.py:272-274
        self.layers = []
        for i in range(args.n_layers):
            p = f"model.layers.{i}."
            cache_shape = (args.max_batch_size, args.max_seq_len, self.n_kv_heads, self.head_dim)
            self.layers.append({
                # stored [out, in]; used transposed (views) - llama3.py:93-95, 133-136
                "wq": weights[p + "self_attn.q_proj.weight"].T,
                "wk": weights[p + "self_attn.k_proj.weight"].T,
                "wv": weights[p + "self_attn.v_proj.weight"].T,
                "wo": weights[p + "self_attn.o_proj.weight"].T,
                "w_up": weights[p + "mlp.up_proj.weight"].T,
                "w_gate": weights[p + "mlp.gate_proj.weight"].T,
                "w_down": weights[p + "mlp.down_proj.weight"].T,
                "norm_in": weights[p + "input_layernorm.weight"],
                "norm_post": weights[p + "post_attention_layernorm.weight"],
                # float64 zero caches, never reset - llama3.py:138-153
                "cache_k": np.zeros(cache_shape),
                "cache_v": np.zeros(cache_shape),
            })
        self.norm_final = weights["model.norm.weight"]
        self.lm_head = weights["lm_head.weight"].T  # llama3.py:281
        if precast:
            for i, layer in enumerate(self.layers):
                for key in ("wo", "w_up", "w_gate", "w_down") + (("wq", "wk", "wv") if i > 0 else ()):
                    layer[key] = np.ascontiguousarray(layer[key], dtype=np.float64)  # C order, as NumPy's own cast temporary
            self.lm_head = np.ascontiguousarray(self.lm_head, dtype=np.float64)

    # ---- one attention call: llama3.py:155-213
    def _attention(self, layer, x, start_pos: int, mask: Optional[np.ndarray], cos, sin):
        B, L, _ = x.shape
        q = (x @ layer["wq"]).reshape(B, L, self.n_heads, self.head_dim)
        k = (x @ layer["wk"]).reshape(B, L, self.n_kv_heads, self.head_dim)
        v = (x @ layer["wv"]).reshape(B, L, self.n_kv_heads, self.head_dim)
        q = rotate_pairs(q, cos, sin)
        k = rotate_pairs(k, cos, sin)
        end = start_pos + L
        layer["cache_k"][:B, start_pos:end] = k  # K cached post-RoPE, V raw (:184-185)
        layer["cache_v"][:B, start_pos:end] = v
        keys = layer["cache_k"][:B, :end]
        vals = layer["cache_v"][:B, :end]
        if self.n_rep != 1:  # llama3.py:79-83: q head h reads kv head h // n_rep
            keys = np.repeat(keys, self.n_rep, axis=2)
            vals = np.repeat(vals, self.n_rep, axis=2)
        q = q.transpose(0, 2, 1, 3)
        keys = keys.transpose(0, 2, 1, 3)
        vals = vals.transpose(0, 2, 1, 3)
        scores = q @ keys.transpose(0, 1, 3, 2) / math.sqrt(self.head_dim)  # :200-202
        if mask is not None:
            scores = scores + mask[None, None, :, :]
        probs = softmax_lastdim(scores)
        ctx = probs @ vals
        ctx = ctx.transpose(0, 2, 1, 3).reshape(B, L, -1)
        return ctx @ layer["wo"]

    # ---- SwiGLU FFN: llama3.py:97-103
    @staticmethod
    def _ffn(layer, x):
        gated = silu(x @ layer["w_gate"])
        up = x @ layer["w_up"]
        return (gated * up) @ layer["w_down"]

    # ---- llama3.py:285-308
    def __call__(self, input_ids: np.ndarray, start_pos: int) -> np.ndarray:
        _, L = input_ids.shape
        h = self.embed[input_ids]
        cos = self.cos[start_pos:start_pos + L]
        sin = self.sin[start_pos:start_pos + L]
        mask = None
        if L > 1:  # [L, start_pos + L]: zeros over the past, strict upper -inf over the chunk
            mask = np.triu(np.full((L, L), float("-inf")), k=1)
            mask = np.concatenate([np.zeros((L, start_pos)), mask], axis=1)
        eps = self.args.norm_eps
        for layer in self.layers:  # llama3.py:239-261
            a = self._attention(layer, rms_norm(h, layer["norm_in"], eps), start_pos, mask, cos, sin)
            z = h + a
            h = z + self._ffn(layer, rms_norm(z, layer["norm_post"], eps))
        h = rms_norm(h, self.norm_final, eps)
        return h[:, [-1], :] @ self.lm_head  # last position only -> [B, 1, VS]

    # ---- llama3.py:310-321
    def generate(self, input_ids: np.ndarray, max_new_tokens: int):
        _, L = input_ids.shape
        next_id = None
        for i, curr_pos in enumerate(range(L, max_new_tokens)):
            if i == 0:
                logits = self(input_ids, 0)
            else:
                logits = self(next_id, curr_pos)  # NB: L + i, not L + i - 1
            next_id = logits[:, -1, :].argmax(-1, keepdims=True)
            yield next_id


def scaled_max_err(got: np.ndarray, want: np.ndarray) -> float:
    """max|got - want| / max|want| - the 'relative error' of the parity bar."""
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-30))
