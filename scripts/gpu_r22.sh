#!/bin/bash
mkdir -p gpurun_out
T=r22
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 200 python scripts/mega_timeline.py llama3-8b 128 8 > gpurun_out/${T}_tl_8b.log 2>&1
timeout 200 python scripts/mega_timeline.py llama3.2-1b 2048 8 > gpurun_out/${T}_tl_1b.log 2>&1
head -40 gpurun_out/${T}_tl_8b.log | tail -20; head -40 gpurun_out/${T}_tl_1b.log | tail -20
timeout 400 python scripts/bench_shapes.py s15m-b1-f32 s15m-b1-bf16 1b 8b-b1 8b-b32 8b-prefill 8b-b1-f32-4l >> gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.log 2>&1; tail -c 2500 gpurun_out/${T}_bench.log
