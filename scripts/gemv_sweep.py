#!/usr/bin/env python
"""Row-streaming GEMV bandwidth at the decode shapes of the 1B / 8B configs (l3_bench_gemv)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200 import _cabi  # noqa: E402

SHAPES = {
    "8b": [("qkv", 6144, 4096), ("wo", 4096, 4096), ("w13", 28672, 4096), ("w2", 4096, 14336), ("lm", 128256, 4096)],
    "1b": [("qkv", 3072, 2048), ("wo", 2048, 2048), ("w13", 16384, 2048), ("w2", 2048, 8192), ("lm", 128256, 2048)],
    "s15m": [("qkv", 864, 288), ("wo", 288, 288), ("w13", 1536, 288), ("w2", 288, 768), ("lm", 32000, 288)],
}

if __name__ == "__main__":
    lib = _cabi.lib()
    bf16 = int(os.environ.get("BF16", "1"))
    rows = int(os.environ.get("ROWS", "1"))
    for model in sys.argv[1:] or ["8b", "1b"]:
        tot_b, tot_ms = 0, 0.0
        for name, n, k in SHAPES[model]:
            ms = C.c_float()
            _cabi.check(lib.l3_bench_gemv(0, n, k, bf16, rows, 200, C.byref(ms)))
            b = n * k * (2 if bf16 else 4)
            tot_b += b
            tot_ms += ms.value
            print(json.dumps({"model": model, "op": name, "N": n, "K": k, "bf16": bf16, "rows": rows,
                              "us": round(ms.value * 1e3, 2), "GBs": round(b / ms.value / 1e6, 1)}), flush=True)
