"""Golden fixtures for the functional surface.  Run in the AUTHORING container:

    python oracle/gen_golden_simple.py

Imports the unmodified `/root/reference/llama3_simple.py` (fp32, MHA only, CORRECT decode
positions pos = L + i - 1, stops at max_seq_len), runs `llama_forward` / `llama_generate` on seeded
weights and stores its outputs in tests/golden/simple_*.npz.  It also checks that stepping the
pinned oracle (`oracle/ref_llama3.py`, float64 activations) through the same position schedule
yields the same tokens, so GPU tests can use either as the checker.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import llama3_simple as ref  # noqa: E402  (the reference's functional implementation)
from config import ModelArgs as RefArgs  # noqa: E402

import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200.synth import make_weights  # noqa: E402
from oracle import ref_llama3 as orc  # noqa: E402
from oracle.gen_golden import weights_digest  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {
    "simple_tiny_mha": (dict(dim=64, n_layers=2, n_heads=4, n_kv_heads=None, vocab_size=96, max_seq_len=32,
                             max_batch_size=3), 160, 21, 3, 5, 40),
    "simple_hd48_mha": (dict(dim=288, n_layers=3, n_heads=6, n_kv_heads=None, vocab_size=1000, max_seq_len=48,
                             max_batch_size=1), 768, 22, 1, 6, 30),
}


def oracle_simple_generate(o, ids, max_new_tokens, max_seq_len):
    """llama3_simple.py:272-285 schedule on the oracle's forward."""
    L = ids.shape[1]
    nxt, cur = None, L
    for i in range(max_new_tokens):
        logits = o(ids, 0) if i == 0 else o(nxt, L + i - 1)
        nxt = logits[:, -1, :].argmax(-1, keepdims=True)
        yield nxt
        cur += 1
        if cur >= max_seq_len:
            break


if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        for name, (fields, hidden, seed, B, L, mnt) in CASES.items():
            args = RefArgs(**fields)
            w = make_weights(args, hidden, seed)
            path = os.path.join(tmp, name + ".npz")
            np.savez(path, **w)
            ids = np.random.default_rng(seed + 1000).integers(0, args.vocab_size, (B, L))
            m = ref.llama_init(path, args)
            logits = ref.llama_forward(m, ids, 0)
            m = ref.llama_init(path, args)
            toks = np.concatenate(list(ref.llama_generate(m, ids, mnt)), axis=1)
            o = orc.OracleLlama(path, args)
            otoks = np.concatenate(list(oracle_simple_generate(o, ids, mnt, args.max_seq_len)), axis=1)
            assert np.array_equal(toks, otoks), name
            out = {"seed": seed, "hidden": hidden, "ids": ids, "max_new_tokens": mnt, "digest": np.array(weights_digest(w)),
                   "logits_prefill": logits, "tokens": toks}
            for k, v in fields.items():
                out["cfg_" + k] = np.array(-1 if v is None else v)
            np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
            print(name, "ok", toks.shape, toks.dtype, logits.dtype, toks[0, :8].tolist())
