#!/bin/bash
mkdir -p gpurun_out
T=r36
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
for env in "X=1" "L3_GEMM_KSPLIT=0" "L3_TF32_BN128=1" "L3_GEMM_KSPLIT=0 L3_TF32_BN128=1"; do
  echo "== $env" >> gpurun_out/${T}_bench.log
  env $env timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/${T}_bench.log
done
python - <<'P'
import json
for line in open('gpurun_out/r36_bench.log'):
    if line.startswith('=='): print(line.strip()); continue
    try:
        d=json.loads(line); print('  value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['per_decode_step_ms'])
    except Exception as e: print('  ??', line[:300])
P
