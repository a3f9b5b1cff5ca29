#!/bin/bash
# mma.sync decode attention: warps per CTA x key split at 8B batch 32 (and the 4-warp form through the op tests)
mkdir -p gpurun_out
L3_ATTN_MMA_NW=4 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -k "attention_decode_bf16 or batched_decode_bf16" 2>&1 | tail -3
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'hbm frac %.3f' % d['decode_hbm_frac'])" | tee -a gpurun_out/r34_ab.log; }
run L3_ATTN_MMA_NW=2 L3_ATTN_TARGET_CTAS=592
run L3_ATTN_MMA_NW=2 L3_ATTN_TARGET_CTAS=256
run L3_ATTN_MMA_NW=4 L3_ATTN_TARGET_CTAS=256
run L3_ATTN_MMA_NW=4 L3_ATTN_TARGET_CTAS=592
run L3_ATTN_MMA_NW=2 L3_ATTN_TARGET_CTAS=888
