#!/bin/bash
mkdir -p gpurun_out
T=r30
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "attention or batched or generate" > gpurun_out/${T}_attn.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_attn.log
tail -5 gpurun_out/${T}_attn.log | cut -c1-300
for env in "X=1" "L3_ATTN_WARP=0"; do
  echo "== $env" >> gpurun_out/${T}_bench.log
  env $env timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/${T}_bench.log
done
python - <<'P'
import json
for line in open('gpurun_out/r30_bench.log'):
    if line.startswith('=='): print(line.strip()); continue
    try:
        d=json.loads(line); print('  value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['per_decode_step_ms'])
    except Exception as e: print('  ??', line[:200])
P
