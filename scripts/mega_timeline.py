#!/usr/bin/env python
"""Per-phase timeline of the persistent decode kernel (L3_MEGA_DBG=1): medians over CTAs, in us."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["L3_MEGA_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa
from llama3_np_b200 import Llama, _cabi
from llama3_np_b200.config import named_config

shape, L, nd = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
args, hidden = named_config(shape, max_batch_size=1, max_seq_len=L + nd + 2, dtype=sys.argv[4] if len(sys.argv) > 4 else "bfloat16")
m = Llama(None, args, hidden_dim=hidden, random_seed=0)
ids = np.random.default_rng(2).integers(3, args.vocab_size, (1, L))
m.generate_all(ids, L + nd)
buf = np.zeros((148, 512), np.uint64)
_cabi.check(m._lib.l3_debug_mega_timeline(m._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), buf.size), m._h)
t = buf.astype(np.float64)
t0 = t[:, 0].min()
ev = {0: "A staged", 1: "A done", 2: "sync1", 3: "attn done", 4: "sync2", 5: "C staged", 6: "C done", 7: "sync3",
      8: "D staged", 9: "D done", 10: "sync4", 11: "E staged", 12: "E done", 13: "sync5"}
for l in (0, 1, 2, args.n_layers - 1):
    if l >= 24:
        continue
    print(f"layer {l}")
    prev = None
    for e, name in ev.items():
        col = t[:, l * 16 + e]
        col = col[col > 0]
        if not len(col):
            continue
        med, lo, hi = (np.median(col) - t0) / 1e3, (col.min() - t0) / 1e3, (col.max() - t0) / 1e3
        d = "" if prev is None else f"  (+{med - prev:6.2f})"
        print(f"  {name:10s} median {med:9.2f}  min {lo:9.2f}  max {hi:9.2f}{d}")
        prev = med
    for mi, name in enumerate(("P qkv", "P wo", "P w13", "P w2")):
        col = t[:, 384 + l * 4 + mi]
        col = col[col > 0]
        print(f"  {name:10s} median {(np.median(col) - t0) / 1e3:9.2f}  min {(col.min() - t0) / 1e3:9.2f}  max {(col.max() - t0) / 1e3:9.2f}")
end = t[:, 383]
print("end", (np.median(end[end > 0]) - t0) / 1e3)
m.close()
