#!/bin/bash
mkdir -p gpurun_out
T=r37
timeout 400 python -m pytest tests/test_mega_gpu.py -m gpu -q -x --timeout 150 -k "batched_persistent" > gpurun_out/${T}_bt.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bt.log
tail -30 gpurun_out/${T}_bt.log | cut -c1-300
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "tcgen05 or batched or generate" > gpurun_out/${T}_tc.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tc.log
tail -4 gpurun_out/${T}_tc.log | cut -c1-300
for env in "X=1" "L3_BATCH_MEGA=0"; do
  echo "== $env" >> gpurun_out/${T}_bench.log
  env $env timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/${T}_bench.log
done
python - <<'P'
import json
for line in open('gpurun_out/r37_bench.log'):
    if line.startswith('=='): print(line.strip()); continue
    try:
        d=json.loads(line); print('  value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['roofline']['per_decode_step_ms'])
    except Exception as e: print('  ??', line[:300])
P
