#!/bin/bash
# 2-GPU call, kept short: TP parity, TP bench B=1/32 (one-shot vs NCCL), 1-GPU same-shape reference
mkdir -p gpurun_out
T=r23
timeout 300 python -m pytest tests/test_tp_gpu.py -x -q --timeout 280 > gpurun_out/${T}_tp_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tp_pytest.log
tail -15 gpurun_out/${T}_tp_pytest.log | cut -c1-300
for env in "L3_TP_ONESHOT=1" "L3_TP_ONESHOT=0"; do
  echo "== $env" >> gpurun_out/${T}_tp_bench.log
  env $env timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_tp.py --layers 32 --batches 1,32 --decode 64 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
done
cat gpurun_out/${T}_tp_bench.log | cut -c1-420
