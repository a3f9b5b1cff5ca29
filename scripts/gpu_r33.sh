#!/bin/bash
mkdir -p gpurun_out
T=r33
timeout 200 python -m pytest tests/test_tp_gpu.py -x -q --timeout 190 > gpurun_out/${T}_tp_pytest.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/${T}_tp_pytest.log
tail -8 gpurun_out/${T}_tp_pytest.log | cut -c1-400
for env in "X=1" "L3_TP_BF16_AR=0"; do
  echo "== $env" >> gpurun_out/${T}_tp_bench.log
  env $env timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_tp.py --layers 32 --batches 1 --prompt 8192 --decode 4 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
done
cat gpurun_out/${T}_tp_bench.log | cut -c1-300
