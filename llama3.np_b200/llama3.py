"""The reference's model surface on a B200: `Llama(model_path, args)`, `__call__(input_ids,
start_pos)`, `generate(input_ids, max_new_tokens)` - same names, argument meaning and state
semantics as `/root/reference/llama3.py:264-321`, with every numerical step running in the
sm_100a kernels behind `include/llama3_b200.h`.

Kept behaviours: logits come back float64 `[B, 1, VS]`; the KV cache is
per-instance state, zero at construction, never reset, addressed by the caller's
`start_pos`; `generate` is a lazy generator yielding `[B, 1]` int64 arrays, runs prefill at
position 0 and decode step i >= 1 at `pos = L + i` (slot L is skipped), and
`max_new_tokens` caps the TOTAL length.  RoPE uses base 10000 whatever `args.rope_theta`
says.  Differences: shape violations raise `ValueError` instead of NumPy broadcast errors;
`args.dtype` picks the arithmetic ("float32": fp32 everywhere, token-identical to the
reference on the tested inputs; "bfloat16": bf16 weights/KV, fp32 accumulation).
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Union

import numpy as np

from . import _cabi
from .config import ModelArgs
from .utils import checkpoint_digest, load_parameters, packed_cache_path


def compute_cos_sin_cache(head_dim: int, max_seq_len: int, base: int = 10000):
    """float64 RoPE tables `[max_seq_len, head_dim // 2]` (reference llama3.py:31-38).
    Computed on the host in float64 and uploaded: fp32 `sincosf` of `t * inv_freq` on the
    device would drift by ~1e-5 at t <= 256 and ~2e-3 rad at t = 32k."""
    inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2)[: head_dim // 2] / head_dim))
    freqs = np.outer(np.arange(max_seq_len), inv_freq)
    return np.cos(freqs), np.sin(freqs)


_DTYPES = {"float32": _cabi.DTYPE_F32, "fp32": _cabi.DTYPE_F32,
           "bfloat16": _cabi.DTYPE_BF16, "bf16": _cabi.DTYPE_BF16}


class Llama:
    def __init__(self, model_path: Union[str, Mapping[str, np.ndarray], None], args: ModelArgs, *,
                 device: int = 0, hidden_dim: Optional[int] = None, random_seed: Optional[int] = None,
                 flags: int = 0, tp_rank: int = 0, tp_world: int = 1, tp_unique_id: Optional[bytes] = None,
                 honor_rope_theta: bool = False, cache_dir=None):
        """`model_path`: an `.npz` in the reference layout (llama3.py:219-235, 269, 280-281) or a
        mapping of the same keys.  Extension for shapes with no checkpoint: `model_path=None`
        with `hidden_dim` and `random_seed` fills the weights on the device.
        Tensor parallel (8B-shaped configs, one process per GPU): `tp_rank`, `tp_world` and the
        128-byte `tp_unique_id` every rank got from rank 0 (`dp.tp_unique_id`); each rank keeps
        its heads / FFN columns / vocabulary rows of the SAME full weight mapping.
        `honor_rope_theta=True` (opt-in, off for oracle parity): build the RoPE tables with
        `args.rope_theta` instead of the reference's hard-coded base 10000 (llama3.py:31, :272-274) -
        needed for real Llama-3 checkpoints (base 500000), see `convert.py`.
        `cache_dir` (extension, only with an `.npz` path): keep the packed device layout of this checkpoint
        (fused q|k|v, interleaved gate/up, this rank's slices, model dtype) in
        `cache_dir/<sha256 of the file>.<dtype>.tp<r>of<w>.l3pack`; the next start streams that file straight
        into the device buffers instead of `np.load` + packing.  A cached load is bit-identical to a fresh pack
        (tests/test_packed_gpu.py); a pack whose digest, shape, dtype, placement or checksums do not match is
        ignored and rewritten."""
        self.args = args
        self._lib = _cabi.lib()
        self._h = C.c_void_p()
        if args.dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {args.dtype!r}; use 'float32' or 'bfloat16'")
        weights = None
        pack_path = pack_digest = None
        self.loaded_from_pack = False
        if cache_dir is not None:
            import os
            if not (isinstance(model_path, (str, bytes)) or hasattr(model_path, "__fspath__")):
                raise ValueError("cache_dir needs model_path to be a checkpoint file (its digest is the cache key)")
            os.makedirs(cache_dir, exist_ok=True)
            pack_digest = checkpoint_digest(model_path)
            pack_path = packed_cache_path(cache_dir, pack_digest, args.dtype, tp_rank, tp_world)
            if os.path.exists(pack_path):
                info, dig = _cabi.L3Config(), C.create_string_buffer(128)
                if self._lib.l3_packed_info(pack_path.encode(), C.byref(info), dig, 128) == _cabi.L3_OK \
                        and dig.value.decode() == pack_digest:
                    hidden_dim = int(info.hidden_dim)
                    self.loaded_from_pack = True
        if self.loaded_from_pack:
            pass
        elif model_path is not None:
            weights = load_parameters(model_path)
            hidden_dim = int(weights["model.layers.0.mlp.up_proj.weight"].shape[0])
        elif hidden_dim is None or random_seed is None:
            raise ValueError("model_path=None needs hidden_dim and random_seed")
        n_kv = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
        cfg = _cabi.L3Config(dim=args.dim, n_layers=args.n_layers, n_heads=args.n_heads, n_kv_heads=n_kv,
                             vocab_size=args.vocab_size, max_seq_len=args.max_seq_len,
                             max_batch_size=args.max_batch_size, hidden_dim=hidden_dim,
                             norm_eps=args.norm_eps, dtype=_DTYPES[args.dtype], device=device,
                             tp_rank=tp_rank, tp_world=tp_world, flags=flags)
        self.hidden_dim = hidden_dim
        self.n_kv_heads = n_kv
        self.head_dim = args.dim // args.n_heads
        if tp_world > 1 and (tp_unique_id is None or len(tp_unique_id) != 128):
            raise ValueError("tp_world > 1 needs the 128-byte tp_unique_id shared by all ranks")
        self.tp_rank, self.tp_world = tp_rank, tp_world
        _cabi.check(self._lib.l3_create(C.byref(cfg), C.byref(self._h)))
        try:
            if tp_world > 1:
                _cabi.check(self._lib.l3_tp_init(self._h, C.c_char_p(tp_unique_id)), self._h)
            if self.loaded_from_pack:
                rc = self._lib.l3_load_packed(self._h, pack_path.encode(), pack_digest.encode())
                if rc == _cabi.L3_EINVAL:  # stale or damaged pack: fall through to the checkpoint itself, then rewrite it
                    self.loaded_from_pack = False
                    weights = load_parameters(model_path)
                    if int(weights["model.layers.0.mlp.up_proj.weight"].shape[0]) != hidden_dim:
                        raise ValueError(f"{pack_path}: header disagrees with the checkpoint; delete it")
                else:
                    _cabi.check(rc, self._h)
            if weights is not None:
                for key in _expected_keys(args):
                    w = weights.get(key) if hasattr(weights, "get") else weights[key]
                    if w is None:
                        raise ValueError(f"missing weight {key!r}")
                    w = np.ascontiguousarray(w, dtype=np.float32)
                    shape = (C.c_int64 * w.ndim)(*w.shape)
                    _cabi.check(self._lib.l3_load_weight(self._h, key.encode(), _cabi.f32p(w), shape, w.ndim), self._h)
                if pack_path is not None:
                    _cabi.check(self._lib.l3_save_packed(self._h, pack_path.encode(), pack_digest.encode()), self._h)
            elif not self.loaded_from_pack:
                _cabi.check(self._lib.l3_fill_random(self._h, random_seed), self._h)
            # RoPE #1 (llama3.py:272-274): rope_theta deliberately not passed, as in the reference
            base = args.rope_theta if honor_rope_theta else 10000
            cos, sin = compute_cos_sin_cache(self.head_dim, args.max_seq_len, base)
            self.freqs_cos, self.freqs_sin = cos, sin
            _cabi.check(self._lib.l3_set_rope_tables(self._h, _cabi.f64p(np.ascontiguousarray(cos)),
                                                     _cabi.f64p(np.ascontiguousarray(sin))), self._h)
            _cabi.check(self._lib.l3_finalize(self._h), self._h)
        except Exception:
            self.close()
            raise

    # ---------------------------------------------------------------- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.l3_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def save_packed(self, path, source_digest: str = "") -> None:
        """Write this model's packed device-layout weights to `path` (see `cache_dir`)."""
        import os
        _cabi.check(self._lib.l3_save_packed(self._h, os.fspath(path).encode(), source_digest.encode()), self._h)

    def reset_cache(self):
        """Zero the KV cache (a fresh reference `Llama` instance starts from zeros)."""
        _cabi.check(self._lib.l3_reset_cache(self._h), self._h)

    # ---------------------------------------------------------------- forward
    def _ids(self, input_ids):
        ids = np.asarray(input_ids)
        if ids.ndim != 2:
            raise ValueError(f"input_ids must be [B, L], got shape {ids.shape}")
        if not np.issubdtype(ids.dtype, np.integer):
            raise ValueError("input_ids must be integers")
        return np.ascontiguousarray(ids, dtype=np.int32)

    def forward_f32(self, input_ids, start_pos: int, want_argmax: bool = False):
        """One step, returning float32 logits `[B, VS]` (and int64 argmax `[B]`)."""
        ids = self._ids(input_ids)
        B, L = ids.shape
        logits = np.empty((B, self.args.vocab_size), dtype=np.float32)
        am = np.empty((B,), dtype=np.int64) if want_argmax else None
        _cabi.check(self._lib.l3_forward(self._h, _cabi.i32p(ids), B, L, int(start_pos), _cabi.f32p(logits),
                                         _cabi.i64p(am) if want_argmax else None), self._h)
        return (logits, am) if want_argmax else logits

    def __call__(self, input_ids, start_pos: int):
        """`[B, L]` ids at `start_pos` -> float64 logits `[B, 1, VS]` of the last position
        (reference llama3.py:285-308)."""
        logits = self.forward_f32(input_ids, start_pos)
        return logits.astype(np.float64)[:, None, :]

    # ---------------------------------------------------------------- generate
    def generate(self, input_ids, max_new_tokens: int):
        """Lazy greedy generator (reference llama3.py:310-321): yields `max_new_tokens - L`
        arrays `[B, 1]` int64; every yielded step is already on the device when it is read."""
        ids = self._ids(input_ids)
        B, L = ids.shape
        n_out = max_new_tokens - L
        if n_out <= 0:
            return
        if max_new_tokens > self.args.max_seq_len:
            raise ValueError(f"max_new_tokens {max_new_tokens} exceeds max_seq_len {self.args.max_seq_len}")
        _cabi.check(self._lib.l3_generate_begin(self._h, _cabi.i32p(ids), B, L), self._h)
        for _ in range(n_out):
            nxt = np.empty((B,), dtype=np.int64)
            _cabi.check(self._lib.l3_generate_next(self._h, _cabi.i64p(nxt)), self._h)
            yield nxt[:, None]

    def generate_all(self, input_ids, max_new_tokens: int) -> np.ndarray:
        """Bulk form of `generate`: the whole greedy loop runs on the device (CUDA-graph
        replays, no per-token host round trip); returns `[B, max_new_tokens - L]` int64."""
        ids = self._ids(input_ids)
        B, L = ids.shape
        n_out = max(0, max_new_tokens - L)
        out = np.empty((B, n_out), dtype=np.int64)
        if n_out:
            _cabi.check(self._lib.l3_generate_greedy(self._h, _cabi.i32p(ids), B, L, int(max_new_tokens),
                                                     _cabi.i64p(out)), self._h)
        return out

    def generate_ragged(self, prompts, max_new_tokens: int, eos_id: Optional[int] = None, schedule: str = "llama3"):
        """Extension: greedy generation for prompts of DIFFERENT lengths in one batch (the reference
        takes equal-length prompts only and leaves EOS to its caller, llama3.py:341-343).  Returns a
        list with `max_new_tokens` new ids per prompt (cut after the first `eos_id`, if given); every
        sequence yields exactly what it would yield alone through `generate` (schedule "llama3", decode
        step i at pos = L + i) or through `llama3_simple.llama_generate` (schedule "simple")."""
        if schedule not in ("llama3", "simple"):
            raise ValueError("schedule must be 'llama3' or 'simple'")
        seqs = [np.ascontiguousarray(np.asarray(p).reshape(-1), dtype=np.int32) for p in prompts]
        if not seqs or any(len(p) == 0 for p in seqs):
            raise ValueError("need at least one non-empty prompt")
        B, lmax = len(seqs), max(len(p) for p in seqs)
        ids = np.zeros((B, lmax), np.int32)
        for b, p in enumerate(seqs):
            ids[b, : len(p)] = p
        lens = np.array([len(p) for p in seqs], np.int32)
        out = np.empty((B, int(max_new_tokens)), np.int64)
        _cabi.check(self._lib.l3_generate_ragged(self._h, _cabi.i32p(ids), _cabi.i32p(lens), B, lmax, int(max_new_tokens),
                                                 0 if schedule == "llama3" else -1, -1 if eos_id is None else int(eos_id),
                                                 _cabi.i64p(out)), self._h)
        res = []
        for b in range(B):
            row = out[b]
            if eos_id is not None and (row == eos_id).any():
                row = row[: int(np.argmax(row == eos_id)) + 1]
            res.append(row.copy())
        return res

    # ---------------------------------------------------------------- state inspection
    def read_cache(self, layer: int):
        """(cache_k, cache_v) of a layer in the reference layout `[max_batch, M, KVHN, HD]`."""
        shape = (self.args.max_batch_size, self.args.max_seq_len, self.n_kv_heads // self.tp_world, self.head_dim)
        k = np.empty(shape, dtype=np.float32)
        v = np.empty(shape, dtype=np.float32)
        _cabi.check(self._lib.l3_read_cache(self._h, layer, _cabi.f32p(k), _cabi.f32p(v)), self._h)
        return k, v

    # ---------------------------------------------------------------- measurement (bench.py)
    def sync(self):
        _cabi.check(self._lib.l3_sync(self._h), self._h)

    def launch_count(self, reset: bool = False) -> int:
        n = C.c_int64()
        _cabi.check(self._lib.l3_launch_count(self._h, C.byref(n), int(reset)), self._h)
        return n.value


def _expected_keys(args: ModelArgs):
    yield "model.embed_tokens.weight"
    for i in range(args.n_layers):
        p = f"model.layers.{i}."
        for s in ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                  "mlp.up_proj", "mlp.gate_proj", "mlp.down_proj", "input_layernorm",
                  "post_attention_layernorm"):
            yield p + s + ".weight"
    yield "model.norm.weight"
    yield "lm_head.weight"
