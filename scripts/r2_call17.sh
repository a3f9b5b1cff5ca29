#!/bin/bash
# 8 GPUs: tensor-parallel decode at TP 8 / TP 4 (batch 1 and 32) with tagged hand-offs, two-level sums and the LL all-reduce
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n scripts/bench_tp.py --batches 1,32 --decode 128 > gpurun_out/r2c17_tp$n.log 2>&1; echo "bench_tp tp$n rc=$?"
  grep '^{' gpurun_out/r2c17_tp$n.log | tee gpurun_out/r2c17_tp$n.jsonl | cut -c1-420; tail -2 gpurun_out/r2c17_tp$n.log | cut -c1-200
done
