#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stack_gpu.py tests/test_mega_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -3
timeout 600 python scripts/stack_determinism.py 256 150 6 24 2>&1 | tail -1
timeout 900 python scripts/stack_sweep.py '{}' '{"L3_LIB_VARIANT":"lpp1"}' '{"L3_LIB_VARIANT":"st16"}' '{"L3_LIB_VARIANT":"st16nb2"}' '{"L3_LIB_VARIANT":"st16nb1"}' '{"L3_LIB_VARIANT":"st16","L3_STACK_PF":128}' '{"L3_LIB_VARIANT":"st16nb2","L3_STACK_PF":128}' | tee gpurun_out/r2c10_sweep.jsonl
for v in "" u2; do
  echo "== decode_mega attention U variant '$v'"
  L3_LIB_VARIANT=$v timeout 600 python scripts/bench_shapes.py 1b 8b-b1 s15m-b1 2>&1 | grep '^{' | tee gpurun_out/r2c10_shapes_${v:-default}.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k in ('config','dtype','decode_tok_s','decode_ms_per_step','decode_hbm_frac','prefill_ms')})"
done
echo "== 8b-b32 tail split"
for t in 0 1; do L3_SWAP_TAILSPLIT=$t timeout 600 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | tee gpurun_out/r2c10_8bb32_tail$t.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k in ('config','decode_tok_s','decode_ms_per_step','decode_hbm_frac')})"; done
