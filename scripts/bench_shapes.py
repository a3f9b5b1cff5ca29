#!/usr/bin/env python
"""Roofline sweep over the other BASELINE.json configs (one JSON line each):

  s15m-b1      stories15M, batch 1, 'I have a dream' cap 256            (configs[0] shape; fp32 + bf16)
  1b           Llama-3.2-1B-shaped, bf16: prefill 2048, then 256 decode (configs[2])
  8b-b1/8b-b32 Llama-3-8B-shaped, bf16: 128-token prompt + 256 decode   (configs[3], single GPU)
  8b-prefill   Llama-3-8B-shaped, bf16: prefill 2048                    (north_star 60 % target)

Random-init weights generated on the device (l3_fill_random); timing = CUDA events on the
library stream; decode tok/s counts generated tokens only; roofline per SURVEY.md 8(d).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200 import Llama, _cabi  # noqa: E402
from llama3_np_b200.config import named_config  # noqa: E402
from llama3_np_b200.synth import param_count  # noqa: E402

HBM, TF = 6545.0, 1672.7
try:
    _p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    HBM, TF = float(_p["hbm_gbs"]), float(_p["bf16_tflops"])
except Exception:
    pass


def timed(m, fn, iters=1):
    ms = C.c_float()
    _cabi.check(m._lib.l3_timer_start(m._h), m._h)
    for _ in range(iters):
        fn()
    _cabi.check(m._lib.l3_timer_stop(m._h, C.byref(ms)), m._h)
    return ms.value / iters


def dev_buf(m, nbytes):
    p = C.c_void_p()
    _cabi.check(m._lib.l3_dev_alloc(m._h, nbytes, C.byref(p)), m._h)
    return p


def run(name, shape, dtype, B, L, n_decode, n_layers=None, flags=0, iters=2):
    over = dict(max_batch_size=B, max_seq_len=L + n_decode + 2, dtype=dtype)
    if n_layers:
        over["n_layers"] = n_layers
    args, hidden = named_config(shape, **over)
    m = Llama(None, args, hidden_dim=hidden, random_seed=0, flags=flags)
    lib, h = m._lib, m._h
    rng = np.random.default_rng(2)
    ids = rng.integers(3, args.vocab_size, (B, L)).astype(np.int32)
    d_ids = dev_buf(m, ids.nbytes)
    _cabi.check(lib.l3_memcpy_h2d(h, d_ids, ids.ctypes.data_as(C.c_void_p), ids.nbytes), h)
    d_out = dev_buf(m, B * max(n_decode, 1) * 8)
    wb = 4 if dtype == "float32" else 2
    nkv = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
    hd = args.dim // args.n_heads
    D, FD, VS, NL = args.dim, hidden, args.vocab_size, args.n_layers
    out = {"config": name, "shape": shape, "dtype": dtype, "B": B, "prompt": L, "decode": n_decode, "n_layers": NL}

    # ---- prefill (logits of the last position + argmax), device-resident ids
    def prefill():
        _cabi.check(lib.l3_forward_dev(h, d_ids, B, L, 0, None, d_out), h)
    prefill()
    m.sync()
    ms_pf = timed(m, prefill, iters)
    flops = 2 * B * L * NL * (2 * D * D + 2 * D * nkv * hd + 3 * D * FD) + B * NL * (4 * L * L * D) // 2 + 2 * B * VS * D
    out["prefill_ms"] = ms_pf
    out["prefill_tok_s"] = B * L / (ms_pf / 1e3)
    out["prefill_tflops"] = flops / (ms_pf / 1e3) / 1e12
    out["prefill_tensor_frac"] = out["prefill_tflops"] / TF

    if n_decode > 1:
        total = L + n_decode

        def gen():
            _cabi.check(lib.l3_generate_greedy_dev(h, d_ids, B, L, total, d_out), h)
        gen()
        m.sync()
        m.launch_count(reset=True)
        ms_gen = timed(m, gen, iters)
        launches = m.launch_count(reset=True) // iters
        ms_dec = (ms_gen - ms_pf) / (n_decode - 1)          # per decode step (first token comes from the prefill)
        params = param_count(args, hidden) - VS * D + D      # weights read per decoded token
        pos_mid = L + n_decode // 2
        bytes_step = params * wb + B * NL * nkv * pos_mid * hd * 2 * wb
        out.update(decode_ms_per_step=ms_dec, decode_tok_s=B / (ms_dec / 1e3), launches_per_generate=launches,
                   weight_bytes_per_step=params * wb, bytes_per_step=bytes_step,
                   decode_hbm_gbs=bytes_step / (ms_dec / 1e3) / 1e9,
                   decode_hbm_frac=bytes_step / (ms_dec / 1e3) / 1e9 / HBM)
    print(json.dumps(out), flush=True)
    lib.l3_dev_free(h, d_ids)
    lib.l3_dev_free(h, d_out)
    m.close()


CONFIGS = {
    "s15m-b1-f32": lambda: run("s15m-b1-f32", "stories15M", "float32", 1, 5, 251, iters=3),
    "s15m-b1-bf16": lambda: run("s15m-b1-bf16", "stories15M", "bfloat16", 1, 5, 251, iters=3),
    "1b": lambda: run("1b", "llama3.2-1b", "bfloat16", 1, 2048, 256),
    "8b-b1": lambda: run("8b-b1", "llama3-8b", "bfloat16", 1, 128, 256),
    "8b-b32": lambda: run("8b-b32", "llama3-8b", "bfloat16", 32, 128, 256),
    "8b-prefill": lambda: run("8b-prefill", "llama3-8b", "bfloat16", 1, 2048, 1),
    "8b-32k": lambda: run("8b-32k", "llama3-8b", "bfloat16", 1, 32768, 1, iters=1),   # configs[4] on ONE GPU
    "1b-prefill": lambda: run("1b-prefill", "llama3.2-1b", "bfloat16", 1, 2048, 1),
    "8b-b1-f32-4l": lambda: run("8b-b1-f32-4l", "llama3-8b", "float32", 1, 128, 64, n_layers=4),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=list(CONFIGS))
    a = ap.parse_args()
    for c in a.configs:
        try:
            CONFIGS[c]()
        except Exception as e:  # keep sweeping
            print(json.dumps({"config": c, "error": f"{type(e).__name__}: {e}"}), flush=True)
