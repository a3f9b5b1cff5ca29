#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-b32 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', 'ms/step %.3f' % d['decode_ms_per_step'], 'tok/s %.0f' % d['decode_tok_s'], 'prefill %.1f' % d['prefill_ms'])"; }
run L3_X=0
run L3_ATTN_TARGET_CTAS=592
run L3_ATTN_TARGET_CTAS=444
run L3_ATTN_TARGET_CTAS=256
run L3_ATTN_TARGET_CTAS=592 L3_LIB_VARIANT=attu4
run L3_ATTN_TARGET_CTAS=256 L3_LIB_VARIANT=attu4
run L3_ATTN_TARGET_CTAS=1184 L3_LIB_VARIANT=attb3
run L3_ATTN_TARGET_CTAS=592 L3_PDL=1
