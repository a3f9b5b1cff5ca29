#!/bin/bash
mkdir -p gpurun_out
T=r32
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "tcgen05 or bf16 or forward or batched" > gpurun_out/${T}_par.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_par.log
tail -5 gpurun_out/${T}_par.log | cut -c1-300
timeout 400 python scripts/bench_shapes.py 1b-prefill 8b-prefill 8b-32k 8b-b32 > gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
