"""The reference's FUNCTIONAL surface on a B200: `llama_init`, `llama_forward`,
`llama_generate` with the names, argument meaning and schedule of
`/root/reference/llama3_simple.py:168-285`, over the same sm_100a kernels as `Llama`.

Differences from `llama3.py` that this module keeps, as the reference's functional file does:
decode step i runs at `pos = L + i - 1` (no skipped cache slot, llama3_simple.py:279), at most
`max_new_tokens` ids are yielded and generation stops when the sequence reaches `max_seq_len`
(:284), ids are int32 `[B, 1]`, and `args.dtype` picks the arithmetic: "float32" (fp32 device mode,
float32 logits) or "float16" / "bfloat16" (bf16 tensor-core mode; logits come back as float16 /
float32).  Extension: GQA checkpoints work (the reference file reshapes k/v with n_heads).

With `import llama3_np_b200.llama3_simple as llama3_simple` the reference's own
`tests/test_llama_implementations.py` compares its `llama3.py` against this GPU path.
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np

from . import _cabi
from .config import ModelArgs
from .llama3 import Llama

_MODES = {"float32": ("float32", np.float32), "float16": ("bfloat16", np.float16),
          "bfloat16": ("bfloat16", np.float32)}


def llama_init(model_path, args: ModelArgs, **kw):
    """Reference `llama_init(model_path, args)` (llama3_simple.py:208-268): returns the model as a
    mapping; the weights, caches and RoPE tables live on the device behind `model["_llama"]`."""
    if args.dtype not in _MODES:
        raise ValueError(f"unsupported dtype {args.dtype!r}; use 'float32' or 'float16'")
    mode, out_dtype = _MODES[args.dtype]
    llama = Llama(model_path, replace(args, dtype=mode), **kw)
    return {"args": args, "dtype": out_dtype, "_llama": llama,
            "freqs_cos": llama.freqs_cos.astype(out_dtype), "freqs_sin": llama.freqs_sin.astype(out_dtype)}


def llama_forward(model, input_ids, start_pos: int):
    """Reference `llama_forward(model, input_ids, start_pos)` (llama3_simple.py:168-203):
    logits `[B, 1, VS]` of the last position in the model's dtype."""
    logits = model["_llama"].forward_f32(input_ids, start_pos)
    return logits.astype(model["dtype"], copy=False)[:, None, :]


def llama_generate(model, input_ids, max_new_tokens: int):
    """Reference `llama_generate` (llama3_simple.py:272-285): lazy generator of int32 `[B, 1]` ids;
    the whole step (forward + argmax) runs on the device, one small read-back per yielded id."""
    llama: Llama = model["_llama"]
    ids = llama._ids(input_ids)
    B, L = ids.shape
    n_out = min(int(max_new_tokens), model["args"].max_seq_len - L)
    if n_out <= 0:
        return
    _cabi.check(llama._lib.l3_generate_begin_ex(llama._h, _cabi.i32p(ids), B, L, -1), llama._h)
    for _ in range(n_out):
        nxt = np.empty((B,), dtype=np.int64)
        _cabi.check(llama._lib.l3_generate_next(llama._h, _cabi.i64p(nxt)), llama._h)
        yield nxt.astype(np.int32)[:, None]


def llama_close(model):
    """Free the device state (extension; the reference relies on garbage collection)."""
    model["_llama"].close()
