"""Pins the oracle and writes the golden fixtures.  Run in the AUTHORING container only:

    python oracle/gen_golden.py

It imports the unmodified reference from /root/reference (read-only, absent on the GPU
box), runs it on seeded synthetic weights, asserts that `oracle/ref_llama3.py` reproduces
every output BIT FOR BIT in this environment, and stores the reference's outputs under
tests/golden/*.npz.  Weights are not stored: fixtures carry (config, seed) and a checksum
of the regenerated weights.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import llama3 as ref  # noqa: E402  (the reference)
from config import ModelArgs as RefArgs  # noqa: E402

import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200.synth import make_weights  # noqa: E402
from oracle import ref_llama3 as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (fields, hidden, seed, batch, prompt_len, cap)
    "tiny_mha": (dict(dim=64, n_layers=2, n_heads=4, n_kv_heads=None, vocab_size=96,
                      max_seq_len=32, max_batch_size=3), 160, 11, 3, 5, 20),
    "tiny_gqa": (dict(dim=128, n_layers=3, n_heads=8, n_kv_heads=2, vocab_size=200,
                      max_seq_len=48, max_batch_size=2), 320, 12, 2, 7, 30),
    "hd48_gqa": (dict(dim=288, n_layers=2, n_heads=6, n_kv_heads=3, vocab_size=512,
                      max_seq_len=64, max_batch_size=4), 768, 13, 4, 6, 40),
    "hd128_gqa": (dict(dim=512, n_layers=2, n_heads=4, n_kv_heads=1, vocab_size=300,
                       max_seq_len=40, max_batch_size=2), 1024, 14, 2, 9, 24),
}


def weights_digest(w):
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k]).tobytes())
    return h.hexdigest()


def same(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b), (what, float(np.max(np.abs(a - b))))


def build(fields, hidden, seed, path):
    args = RefArgs(**fields)
    w = make_weights(args, hidden, seed)
    np.savez(path, **w)
    return args, w


def model_case(name, tmp):
    fields, hidden, seed, B, L, cap = CASES[name]
    path = os.path.join(tmp, name + ".npz")
    args, w = build(fields, hidden, seed, path)
    rng = np.random.default_rng(seed + 1000)
    ids = rng.integers(0, args.vocab_size, (B, L))
    out = {"seed": seed, "hidden": hidden, "ids": ids, "cap": cap,
           "digest": np.array(weights_digest(w))}
    for k, v in fields.items():
        out["cfg_" + k] = np.array(-1 if v is None else v)

    # (i) prefill, decode at L, decode at L+1 on one instance (cache carried)
    r, o = ref.Llama(path, args), orc.OracleLlama(path, args)
    nxt = rng.integers(0, args.vocab_size, (B, 1))
    nxt2 = rng.integers(0, args.vocab_size, (B, 1))
    for tag, (x, pos) in {"prefill": (ids, 0), "decode0": (nxt, L), "decode1": (nxt2, L + 1)}.items():
        a, b = r(x, pos), o(x, pos)
        same(a, b, (name, tag))
        out["logits_" + tag] = a
    out["nxt"], out["nxt2"] = nxt, nxt2

    # (ii) chunked prefill on a fresh instance: first 3, then the rest at start_pos=3
    r, o = ref.Llama(path, args), orc.OracleLlama(path, args)
    same(r(ids[:, :3], 0), o(ids[:, :3], 0), (name, "chunk0"))
    a, b = r(ids[:, 3:], 3), o(ids[:, 3:], 3)
    same(a, b, (name, "chunk1"))
    out["logits_chunked"] = a

    # (iii) generate (position quirk + total-length cap), B rows at once, then a second
    # generate on the SAME instance with a shorter prompt (stale-cache behaviour)
    r, o = ref.Llama(path, args), orc.OracleLlama(path, args)
    ta = np.concatenate(list(r.generate(ids, cap)), axis=1)
    tb = np.concatenate(list(o.generate(ids, cap)), axis=1)
    same(ta, tb, (name, "generate"))
    assert ta.shape == (B, cap - L)
    out["tokens"] = ta
    ids2 = ids[:, : L - 2]
    ta2 = np.concatenate(list(r.generate(ids2, cap - 3)), axis=1)
    tb2 = np.concatenate(list(o.generate(ids2, cap - 3)), axis=1)
    same(ta2, tb2, (name, "generate2"))
    out["tokens_second"] = ta2
    # final cache of layer 0 (reference layout [maxB, M, KVHN, HD])
    same(r.layers[0].attention.cache_k, o.layers[0]["cache_k"], (name, "cache_k"))
    out["cache_k0"] = r.layers[0].attention.cache_k
    out["cache_v0"] = r.layers[0].attention.cache_v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: ok  tokens[0,:8]={ta[0, :8].tolist()}")


def stories_case(tmp):
    """BASELINE.json configs[0] on random-init weights in the stories15M layout."""
    fields = dict(dim=288, n_layers=6, n_heads=6, n_kv_heads=None, vocab_size=32000,
                  max_seq_len=256, max_batch_size=1)
    path = os.path.join(tmp, "s15m.npz")
    args, w = build(fields, 768, 0, path)
    ids = np.array([[1, 76, 505, 263, 12561]])  # "I have a dream" via the reference tokenizer
    r, o = ref.Llama(path, args), orc.OracleLlama(path, args)
    a, b = r(ids, 0), o(ids, 0)
    same(a, b, "s15m prefill")
    r, o = ref.Llama(path, args), orc.OracleLlama(path, args)
    ta = np.concatenate(list(r.generate(ids, 50)), axis=1)
    tb = np.concatenate(list(o.generate(ids, 50)), axis=1)
    same(ta, tb, "s15m generate")
    assert ta.shape == (1, 45)
    srt = np.sort(a[0, 0])[::-1]
    np.savez_compressed(os.path.join(GOLD, "stories15m_c1.npz"), seed=0, hidden=768, ids=ids,
                        cap=50, digest=np.array(weights_digest(w)),
                        logits_prefill=a.astype(np.float32), top5=np.argsort(-a[0, 0])[:5],
                        top_gap=srt[0] - srt[1], tokens=ta)
    print("stories15m_c1: ok", ta[0, :10].tolist(), "gap", srt[0] - srt[1])


def op_case():
    """Per-op reference outputs (functions the reference's own unit tests exercise)."""
    rng = np.random.default_rng(7)
    out = {}
    x = rng.standard_normal((1, 6, 8, 8)).astype(np.float32)
    out["softmax_in"], out["softmax_out"] = x, ref.softmax(x)
    same(out["softmax_out"], orc.softmax_lastdim(x), "softmax")
    x = rng.standard_normal((1, 8, 288)).astype(np.float32)
    out["silu_in"], out["silu_out"] = x, ref.silu(x)
    same(out["silu_out"], orc.silu(x), "silu")
    c, s = ref.compute_cos_sin_cache(48, 256)
    c2, s2 = orc.rope_tables(48, 256)
    same(c, c2, "cos"); same(s, s2, "sin")
    out["cos48"], out["sin48"] = c, s
    xq = rng.standard_normal((2, 8, 6, 48)).astype(np.float32)
    xk = rng.standard_normal((2, 8, 3, 48)).astype(np.float32)
    rq, rk = ref.apply_rotary_emb(xq, xk, c[4:12], s[4:12])
    same(rq, orc.rotate_pairs(xq, c[4:12], s[4:12]), "rope q")
    same(rk, orc.rotate_pairs(xk, c[4:12], s[4:12]), "rope k")
    out.update(rope_q_in=xq, rope_k_in=xk, rope_q_out=rq, rope_k_out=rk)
    w = rng.standard_normal(288).astype(np.float32)
    x = rng.standard_normal((1, 8, 288)).astype(np.float32)
    y = ref.RMSNorm(w, 1e-6)(x)
    same(y, orc.rms_norm(x, w, 1e-6), "rmsnorm")
    out.update(rms_in=x, rms_w=w, rms_out=y)
    up = (0.05 * rng.standard_normal((160, 96))).astype(np.float32)
    gate = (0.05 * rng.standard_normal((160, 96))).astype(np.float32)
    down = (0.05 * rng.standard_normal((96, 160))).astype(np.float32)
    x = rng.standard_normal((2, 5, 96)).astype(np.float32)
    y = ref.FeedForward(up, gate, down)(x)
    same(y, orc.OracleLlama._ffn({"w_up": up.T, "w_gate": gate.T, "w_down": down.T}, x), "ffn")
    out.update(ffn_in=x, ffn_up=up, ffn_gate=gate, ffn_down=down, ffn_out=y)
    np.savez_compressed(os.path.join(GOLD, "ops.npz"), **out)
    print("ops: ok")


if __name__ == "__main__":
    import tempfile
    os.makedirs(GOLD, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        op_case()
        for name in CASES:
            model_case(name, tmp)
        stories_case(tmp)
    print("oracle pinned: bit-exact against /root/reference on all cases")
