#!/bin/bash
mkdir -p gpurun_out
T=r25
timeout 170 python -m pytest tests/test_tp_gpu.py -x -q --timeout 160 > gpurun_out/${T}_tp_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tp_pytest.log
tail -12 gpurun_out/${T}_tp_pytest.log | cut -c1-400
for env in "L3_TP_ONESHOT=1" "L3_TP_ONESHOT=0"; do
  echo "== $env" >> gpurun_out/${T}_tp_bench.log
  env $env timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_tp.py --layers 32 --batches 1,32 --decode 64 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
done
cat gpurun_out/${T}_tp_bench.log | cut -c1-420
