#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -x -q --timeout 800 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_tp.py --batches 1,32 --decode 128 > gpurun_out/r2c14_tp2.log 2>&1; echo "bench_tp rc=$?"; grep '^{' gpurun_out/r2c14_tp2.log | tee gpurun_out/r2c14_tp2.jsonl | cut -c1-420
