#!/bin/bash
mkdir -p gpurun_out
T=r21
timeout 300 python -m pytest tests/test_mega_gpu.py -m gpu -q --timeout 90 > gpurun_out/${T}_mega.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_mega.log
tail -5 gpurun_out/${T}_mega.log
timeout 200 python scripts/mega_timeline.py llama3-8b 128 8 > gpurun_out/${T}_tl_8b.log 2>&1
timeout 200 python scripts/mega_timeline.py llama3.2-1b 2048 8 > gpurun_out/${T}_tl_1b.log 2>&1
head -40 gpurun_out/${T}_tl_8b.log | tail -20; head -40 gpurun_out/${T}_tl_1b.log | tail -20
timeout 300 python scripts/bench_shapes.py s15m-b1-f32 1b 8b-b1 >> gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
