#!/bin/bash
# 8 GPUs: tensor-parallel batch-1 decode latency at TP 8 and TP 4 with the flag-in-data exchange
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/bench_tp.py --batches 1 --decode 128 > gpurun_out/r2c12_tp$n.log 2>&1; echo "bench_tp tp$n rc=$?"
  grep '^{' gpurun_out/r2c12_tp$n.log | tee gpurun_out/r2c12_tp$n.jsonl | cut -c1-600; tail -2 gpurun_out/r2c12_tp$n.log | cut -c1-300
done
