#!/bin/bash
# 2-GPU call: TP parity test, TP bench (NCCL-only vs one-shot), then single-GPU regression + shapes
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r13_topo.log 2>&1
timeout 600 python -m pytest tests/test_tp_gpu.py -x -q --timeout 600 > gpurun_out/r13_tp_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r13_tp_pytest.log
tail -30 gpurun_out/r13_tp_pytest.log
for env in "L3_TP_ONESHOT=1" "L3_TP_ONESHOT=0"; do
  echo "== $env" >> gpurun_out/r13_tp_bench.log
  env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_tp.py --layers 32 --batches 1,32 --decode 64 >> gpurun_out/r13_tp_bench.log 2>&1
done
tail -12 gpurun_out/r13_tp_bench.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/r13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r13_pytest.log
tail -5 gpurun_out/r13_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python scripts/bench_shapes.py s15m-b1-f32 1b 8b-b1 > gpurun_out/r13_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/r13_shapes.log
CUDA_VISIBLE_DEVICES=0 timeout 100 python scripts/gemv_sweep.py 8b 1b > gpurun_out/r13_gemv.log 2>&1; tail -10 gpurun_out/r13_gemv.log
