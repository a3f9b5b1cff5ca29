#!/bin/bash
mkdir -p gpurun_out
T=r27
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "swapped" > gpurun_out/${T}_swap.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_swap.log
tail -12 gpurun_out/${T}_swap.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-300
for env in "X=1" "L3_GEMM_SWAP=0"; do
  echo "== $env" >> gpurun_out/${T}_shapes.log
  env $env timeout 300 python scripts/bench_shapes.py 8b-b32 >> gpurun_out/${T}_shapes.log 2>&1
done
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.log 2>&1; python - <<'P'
import json
d=json.loads(open('gpurun_out/r27_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['per_decode_step_ms'], d['roofline']['traffic'])
P
