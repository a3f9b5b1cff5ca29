"""Determinism detector for the cluster-resident decode kernel: the KV caches after a short generate must be
bit-identical run to run (every reduction has a fixed order).  Prints how many runs deviate from the majority."""
import os, sys, hashlib, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import llama3_np_b200  # noqa
from llama3_np_b200 import Llama, ModelArgs, _cabi
from llama3_np_b200.synth import make_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
NR = int(sys.argv[2]) if len(sys.argv) > 2 else 200
NL = int(sys.argv[3]) if len(sys.argv) > 3 else 6
TOK = int(sys.argv[4]) if len(sys.argv) > 4 else 16
args = ModelArgs(dim=288, n_layers=NL, n_heads=6, n_kv_heads=6, vocab_size=2000, max_seq_len=40, max_batch_size=B)
w = make_weights(args, 768, seed=12)
ids = np.random.default_rng(6).integers(3, 2000, (B, 5))
m = Llama(w, args)
hs = []
for r in range(NR):
    m.reset_cache()
    m.generate_all(ids, TOK)
    h = hashlib.sha1()
    k, v = m.read_cache(NL - 1)   # the last layer's rows depend on everything before them
    h.update(k.tobytes()); h.update(v.tobytes())
    hs.append(h.hexdigest())
m.close()
c = collections.Counter(hs)
maj = c.most_common(1)[0][1]
print(f"B={B} runs={NR}: {NR - maj} deviate from the majority ({len(c)} distinct results)")
