#!/bin/bash
mkdir -p gpurun_out
T=r19
timeout 300 python -m pytest tests/test_mega_gpu.py -m gpu -q --timeout 90 > gpurun_out/${T}_mega.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_mega.log
tail -5 gpurun_out/${T}_mega.log
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 90 -k "attention or generate or stories or batched" > gpurun_out/${T}_par.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_par.log
tail -5 gpurun_out/${T}_par.log
for env in "L3_MEGA_AHEAD=12" "L3_MEGA_AHEAD=0" "L3_MEGA_AHEAD=30" "L3_MEGA=0"; do
  echo "== $env" >> gpurun_out/${T}_shapes.log
  env $env timeout 300 python scripts/bench_shapes.py s15m-b1-f32 1b 8b-b1 >> gpurun_out/${T}_shapes.log 2>&1
done
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
