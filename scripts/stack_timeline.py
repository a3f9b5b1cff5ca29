#!/usr/bin/env python
"""Phase timeline of the cluster-resident decode kernel (decode_stack.cu) on the headline shape: %globaltimer
stamps of thread 0 of every CTA for the LAST decode step of a generate to `--len` tokens, averaged over CTAs.

  L3_STACK_DBG=1 is set here; the dumps it enables add a few stores per phase, so absolute times run slightly high."""
import argparse, ctypes as C, os, sys
os.environ["L3_STACK_DBG"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa
from llama3_np_b200 import Llama, _cabi
from llama3_np_b200.config import named_config
from llama3_np_b200.synth import make_weights

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--len", type=int, default=134)
a = ap.parse_args()
args, hidden = named_config("stories15M", max_batch_size=a.B)
w = make_weights(args, hidden, seed=0)
m = Llama(w, args)
ids = np.random.default_rng(1).integers(3, 32000, (a.B, 8)).astype(np.int32)
m.generate_all(ids, a.len)
ncl = (a.B + 11) // 12
st = np.zeros((ncl * 8, 128), np.uint64)
_cabi.check(m._lib.l3_debug_stack(m._h, 0, st.ctypes.data_as(C.c_void_p), st.nbytes), m._h)
m.close()
st = st[: ncl * 6].astype(np.int64)
t0 = st[:, 63].min()
names = ["A: norm + QKV + rope", "attention + merge", "B: Wo + push", "exchange 1", "C: norm + gate/up + silu", "D: Wdown + push", "exchange 2"]
print(f"B={a.B} position {a.len - 1}: kernel span {(st[:, 62].max() - t0) / 1e3:.1f} us; CTA start spread {(st[:, 63].max() - t0) / 1e3:.1f} us; "
      f"setup + embedding {np.mean(st[:, 0] - st[:, 63]) / 1e3:.2f} us")
tot = np.zeros(7)
for l in range(6):
    prev = st[:, 0] if l == 0 else st[:, 7 + (l - 1) * 10]
    row = []
    for e in range(7):
        cur = st[:, 1 + e + l * 10]
        row.append(np.mean(cur - prev) / 1e3)
        prev = cur
    tot += np.array(row)
    print(f"layer {l}: " + "  ".join(f"{n.split(':')[0][:10]} {v:5.2f}" for n, v in zip(names, row)) + f"   = {sum(row):.2f} us")
print("sum over layers (us): " + "; ".join(f"{n} {v:.1f}" for n, v in zip(names, tot)) + f"; total {tot.sum():.1f}")
print(f"final norm + store {np.mean(st[:, 62] - st[:, 57]) / 1e3:.2f} us; per-CTA busy span mean {(np.mean(st[:, 62] - st[:, 63])) / 1e3:.1f} us")
w = st[:, [8 + 10 * l for l in range(5)]].mean() / 1.965e3, st[:, [9 + 10 * l for l in range(5)]].mean() / 1.965e3, st[:, [10 + 10 * l for l in range(5)]].mean() / 1.965e3
print(f"thread 0 waiting for ring data per layer (us at 1965 MHz): A + B {w[0]:.2f}, C {w[1]:.2f}, D {w[2]:.2f}")
mc = st[:, 58:62].mean(axis=0) / 1.965e3
print(f"layer 2, thread 0: cycles from stage-ready to last FFMA, summed over the slabs (us): A {mc[0]:.2f}, B {mc[1]:.2f}, C {mc[2]:.2f}, D {mc[3]:.2f}")
for w in range(8):
    d = st[:, 64 + 4 * w: 68 + 4 * w]
    print(f"  layer 2 attention, warp {w}: CTA units {d[:, 0].mean():.1f}  waiting {d[:, 1].mean() / 1.965e3:.2f} us  math {d[:, 2].mean() / 1.965e3:.2f} us  done {np.mean(d[:, 3] - st[:, 21]) / 1e3:.2f} us after the phase began")
print(f"  layer 2: attention phase (stamp 21 -> 22) {np.mean(st[:, 22] - st[:, 21]) / 1e3:.2f} us")

# sub-phase stamps of layer 2 (slots 96 ..): 0 after norm A, 1 after QKV GEMM, 8 attention units drained, 7 after Wo GEMM,
# 2 after norm C, 3 after gate/up GEMM, 4 after Wdown GEMM, 5 partial sums arrived, 6 reduce + gather issued
x = st[:, 96:112]
def d(a_, b_):
    return np.mean(b_ - a_) / 1e3
L2 = 20
print("layer 2 sub-phases (us): "
      f"A: norm {d(st[:, 17], x[:, 0]):.2f} | QKV gemm {d(x[:, 0], x[:, 1]):.2f} | rope + cache {d(x[:, 1], st[:, 21]):.2f};  "
      f"attention: units {d(st[:, 21], x[:, 8]):.2f} | merge {d(x[:, 8], st[:, 22]):.2f};  "
      f"B: Wo gemm {d(st[:, 22], x[:, 7]):.2f} | push {d(x[:, 7], st[:, 23]):.2f};  "
      f"C: norm {d(st[:, 24], x[:, 2]):.2f} | gate/up gemm {d(x[:, 2], x[:, 3]):.2f} | silu {d(x[:, 3], st[:, 25]):.2f};  "
      f"D: Wdown gemm {d(st[:, 25], x[:, 4]):.2f} | push {d(x[:, 4], st[:, 26]):.2f};  "
      f"exchange 2: partials arrive {d(st[:, 26], x[:, 5]):.2f} | reduce + gather stores {d(x[:, 5], x[:, 6]):.2f} | gather arrives {d(x[:, 6], st[:, 27]):.2f}")
