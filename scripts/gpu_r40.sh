#!/bin/bash
mkdir -p gpurun_out
T=r40
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "tcgen05_prefill or long_prompt or bf16" > gpurun_out/${T}_attn.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_attn.log
tail -12 gpurun_out/${T}_attn.log | cut -c1-300
timeout 400 python scripts/bench_shapes.py 1b-prefill 8b-prefill 8b-32k > gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
