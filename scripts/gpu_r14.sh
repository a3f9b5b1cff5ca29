#!/bin/bash
mkdir -p gpurun_out
T=r14
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x --timeout 90 -k "op_linear or op_attention or op_rms or op_rope or op_swiglu or op_argmax" > gpurun_out/${T}_ops.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_ops.log
tail -12 gpurun_out/${T}_ops.log
timeout 400 python -m pytest tests/test_mega_gpu.py -m gpu -q --timeout 90 > gpurun_out/${T}_mega.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_mega.log
tail -25 gpurun_out/${T}_mega.log
timeout 600 python -m pytest tests -m gpu -q --timeout 120 --deselect tests/test_mega_gpu.py > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -25 gpurun_out/${T}_pytest.log
timeout 400 python scripts/bench_shapes.py s15m-b1-f32 1b 8b-b1 8b-prefill > gpurun_out/${T}_shapes.log 2>&1
python scripts/show_shapes.py gpurun_out/${T}_shapes.log
