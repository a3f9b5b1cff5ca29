#!/usr/bin/env python
"""Cycles per tcgen05.mma (M = 128, K = 32 bytes) by kind, width N and accumulator rotation - the tensor pipe
alone, operands resident in shared memory (csrc/mma_probe.cu).  One JSON line per point.

  python scripts/mma_cost.py            # one CTA on an idle chip
  python scripts/mma_cost.py --ctas 148 # every SM at once (power / clock effects)

Reading: `total` / `peak` is the fraction of the dense tensor rate the MMA stream reaches; a flat `total` over N
means narrow tiles waste the pipe."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ctas", type=int, default=1)
    ap.add_argument("--iters", type=int, default=2048)
    a = ap.parse_args()
    lib = _cabi.lib()
    out = (C.c_double * 2)()
    for kind, name, flop_per_clk in ((0, "bf16", 8192.0), (1, "tf32", 4096.0)):  # dense per-SM peak at 1 MMA stream
        for n in (16, 32, 64, 128, 256):
            for nacc in (1, 2, 4):
                if nacc * n > 512:
                    continue
                rc = lib.l3_probe_mma(0, kind, n, nacc, a.iters, a.ctas, out)
                if rc != 0:
                    print(json.dumps({"kind": name, "n": n, "nacc": nacc, "error": rc}))
                    continue
                k = 16 if kind == 0 else 8
                ideal = 2.0 * 128 * n * k / flop_per_clk
                print(json.dumps({"kind": name, "n": n, "nacc": nacc, "ctas": a.ctas, "issue_cycles": round(out[0], 1),
                                  "total_cycles": round(out[1], 1), "ideal_cycles_at_nominal_peak": round(ideal, 1),
                                  "frac_of_nominal": round(ideal / out[1], 3)}))


if __name__ == "__main__":
    main()
