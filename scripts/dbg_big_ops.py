"""Isolate a failing kernel at 1B-class shapes: one op per process (argv[1])."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa
from llama3_np_b200 import _cabi, Llama
from llama3_np_b200.config import named_config
lib = _cabi.lib()
what = sys.argv[1]
rng = np.random.default_rng(0)
if what.startswith("gemm"):
    rows, n, k = {"gemm_w13": (2048, 16384, 2048), "gemm_qkv": (2048, 3072, 2048), "gemm_wo": (2048, 2048, 2048),
                  "gemm_small": (512, 4096, 512)}[what]
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    rc = lib.l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, 3, 1, _cabi.f32p(out))
    import torch
    xb = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    wb = torch.from_numpy(w).to(torch.bfloat16).float().numpy()
    want = xb @ wb.T
    print(what, "rc", rc, "err", float(np.abs(out - want).max() / np.abs(want).max()))
elif what.startswith("attn"):
    B, L, HN, KVHN, HD = {"attn_1b": (1, 2048, 32, 8, 64), "attn_8b": (1, 1024, 32, 8, 128), "attn_mid": (1, 1024, 8, 2, 64)}[what]
    q = rng.standard_normal((B, L, HN, HD)).astype(np.float32)
    k = rng.standard_normal((B, L, KVHN, HD)).astype(np.float32)
    v = rng.standard_normal((B, L, KVHN, HD)).astype(np.float32)
    out = np.empty((B, L, HN * HD), np.float32)
    rc = lib.l3_op_attention(0, _cabi.f32p(q), _cabi.f32p(k), _cabi.f32p(v), B, L, HN, KVHN, HD, 0, 2, 0, _cabi.f32p(out))
    print(what, "rc", rc, "finite", bool(np.isfinite(out).all()), "absmax", float(np.abs(out).max()))
elif what.startswith("model"):
    nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    M = int(sys.argv[4]) if len(sys.argv) > 4 else L + 66
    am = len(sys.argv) > 5 and sys.argv[5] == "argmax"
    args, hidden = named_config("llama3.2-1b", max_batch_size=1, max_seq_len=M, dtype="bfloat16", n_layers=nl)
    m = Llama(None, args, hidden_dim=hidden, random_seed=0)
    ids = rng.integers(3, args.vocab_size, (1, L))
    lg = m.forward_f32(ids, 0, want_argmax=am)
    lg = lg[0] if am else lg
    print(what, sys.argv[2:], "prefill ok", bool(np.isfinite(lg).all()), flush=True)
    toks = m.generate_all(ids[:, :64], 64 + 32)
    print(what, "generate ok", toks[0, :8])
    m.close()
