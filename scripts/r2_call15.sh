#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -x -q --timeout 800 2>&1 | tail -3
for ll in 1 0; do
L3_TP_LL=$ll timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ll scripts/bench_tp.py --batches 32 --decode 128 > gpurun_out/r2c15_tp2_ll$ll.log 2>&1; echo "bench_tp ll=$ll rc=$?"; grep '^{' gpurun_out/r2c15_tp2_ll$ll.log | tee gpurun_out/r2c15_tp2_ll$ll.jsonl | cut -c1-420
done
