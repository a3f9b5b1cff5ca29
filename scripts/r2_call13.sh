#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_mega_gpu.py tests/test_parity2_gpu.py tests/test_packed_gpu.py -m gpu -x -q --timeout 900 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python scripts/bench_shapes.py 1b 8b-b1 s15m-b1 2>&1 | grep '^{' | tee gpurun_out/r2c13_shapes.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in d if k in ('config','dtype','decode_tok_s','decode_ms_per_step','decode_hbm_frac','prefill_ms')})"
echo "== 1b, context 2048"; timeout 300 python scripts/mega_timeline.py llama3.2-1b 2048 8 > gpurun_out/r2c13_mega_1b_2048.txt 2>&1; sed -n 20,38p gpurun_out/r2c13_mega_1b_2048.txt
