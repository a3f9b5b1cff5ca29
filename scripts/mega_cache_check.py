#!/usr/bin/env python
"""Which path deviates, where: KV caches of the persistent batch-1 kernel and of the kernel-per-projection path after the
same greedy generate, against the oracle's caches, repeated N times (a hand-off race would show as a rare wrong row)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama3_np_b200  # noqa
from llama3_np_b200 import Llama, ModelArgs, _cabi
from llama3_np_b200.synth import make_weights
from oracle import ref_llama3 as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d, nl, hn, kv, hid, vs, msl = (1024, 2, 8, 2, 4608, 1024, 96)
args = ModelArgs(dim=d, n_layers=nl, n_heads=hn, n_kv_heads=kv, vocab_size=vs, max_seq_len=msl, max_batch_size=1)
w = make_weights(args, hid, seed=11)
ids = np.random.default_rng(3).integers(3, vs, (1, 6))
o = orc.OracleLlama(w, args)
want = np.concatenate(list(o.generate(ids, msl)), axis=1)
ref = [(o.layers[l]["cache_k"][:1].astype(np.float32), o.layers[l]["cache_v"][:1].astype(np.float32)) for l in range(nl)]
for name, flags in (("mega", 0), ("per-kernel", _cabi.FLAG_NO_MEGA)):
    m = Llama(w, args, flags=flags)
    bad = 0
    for it in range(N):
        m.reset_cache()
        got = m.generate_all(ids, msl) if it % 2 == 0 else np.concatenate(list(m.generate(ids, msl)), axis=1)
        tok_ok = np.array_equal(got, want)
        for l in range(nl):
            k, v = m.read_cache(l)
            for nm, a, b in (("k", k, ref[l][0]), ("v", v, ref[l][1])):
                diff = np.abs(a - b).max(axis=(0, 2, 3))   # per position
                rows = np.nonzero(diff > 2e-5)[0]
                if len(rows) or not tok_ok:
                    bad += 1
                    print(f"{name} iter {it} ({'bulk' if it % 2 == 0 else 'lazy'}) layer {l} {nm}: tokens_ok={tok_ok} positions {rows.tolist()} max {diff.max():.3g}")
    print(f"{name}: {bad} deviating (iteration, layer, tensor) triples in {N} iterations")
    m.close()
