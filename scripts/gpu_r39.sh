#!/bin/bash
mkdir -p gpurun_out
T=r39
for env in "X=1" "L3_TP_BF16_AR=0"; do
  echo "== tp 8, 32k prefill, $env" >> gpurun_out/${T}_tp_bench.log
  env $env timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/bench_tp.py --layers 32 --batches 1 --prompt 32768 --decode 4 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
done
cat gpurun_out/${T}_tp_bench.log | cut -c1-330
