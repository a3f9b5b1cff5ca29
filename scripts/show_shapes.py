"""Pretty-print bench_shapes.py JSON lines."""
import json
import sys

for f in sys.argv[1:]:
    print(f)
    for line in open(f):
        try:
            d = json.loads(line)
        except Exception:
            print("   ", line[:160].rstrip())
            continue
        if "error" in d:
            print(f"  {d['config']:14s} ERROR {d['error'][:120]}")
            continue
        s = f"  {d['config']:14s} prefill {d['prefill_ms']:8.2f} ms  {d['prefill_tok_s']:9.0f} tok/s  tensor {100 * d['prefill_tensor_frac']:5.1f}%"
        if "decode_tok_s" in d:
            s += f" | decode {d['decode_ms_per_step']:7.3f} ms/step {d['decode_tok_s']:8.0f} tok/s  HBM {100 * d['decode_hbm_frac']:5.1f}%"
        print(s)
