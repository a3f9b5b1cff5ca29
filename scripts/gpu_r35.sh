#!/bin/bash
mkdir -p gpurun_out
T=r35
L3_PDL=1 L3_GEMM_MAXSTAGES=3 timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${T}_pytest_pdl.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_pdl.log
tail -6 gpurun_out/${T}_pytest_pdl.log | cut -c1-300
for env in "X=1" "L3_PDL=1 L3_GEMM_MAXSTAGES=3"; do
  echo "== $env" >> gpurun_out/${T}_shapes.log
  env $env timeout 300 python scripts/bench_shapes.py 8b-b32 8b-prefill 1b >> gpurun_out/${T}_shapes.log 2>&1
done
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
