#!/bin/bash
# 8-GPU call: TP parity (2 ranks), then 8B-shaped TP decode / prefill at TP 2, 4, 8 and the 32k prefill at TP 8.
mkdir -p gpurun_out
T=r26
timeout 170 python -m pytest tests/test_tp_gpu.py -x -q --timeout 160 > gpurun_out/${T}_tp_pytest.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/${T}_tp_pytest.log
tail -8 gpurun_out/${T}_tp_pytest.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 1; fi
for n in 8 4 2; do
  echo "== tp $n" >> gpurun_out/${T}_tp_bench.log
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scripts/bench_tp.py --layers 32 --batches 1,32 --decode 64 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
done
echo "== tp 8, 32k prefill" >> gpurun_out/${T}_tp_bench.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/bench_tp.py --layers 32 --batches 1 --prompt 32768 --decode 4 2>&1 | grep -E "config|Error|error" >> gpurun_out/${T}_tp_bench.log
cat gpurun_out/${T}_tp_bench.log | cut -c1-450
