#!/bin/bash
# RoPE table loads hoisted in both GEMM epilogues, one-pass RMSNorm, staged decode attention: full suite, then new vs base
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6 | tee gpurun_out/r30_pytest.log
run() { env "$@" timeout 300 python scripts/bench_shapes.py 8b-prefill 1b 8b-b32 8b-b1 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$*', d['config'], 'prefill ms %.2f' % d['prefill_ms'], 'tensor frac %.3f' % d['prefill_tensor_frac'], ('decode ms %.3f hbm frac %.3f' % (d['decode_ms_per_step'], d['decode_hbm_frac'])) if 'decode_ms_per_step' in d else '')" | tee -a gpurun_out/r30_ab.log; }
run L3_X=0
run L3_LIB_VARIANT=base
run L3_PDL=1
