#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -x -q --timeout 800 2>&1 | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 > gpurun_out/r2c21_bench_n2.log 2> gpurun_out/r2c21_bench_n2.err ) 2>&1 | grep real
tail -1 gpurun_out/r2c21_bench_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
for r in d.get('tp', []):
    print({k: r.get(k) for k in ('name','tp','value','ms_per_step','ranks_agree_on_tokens','error')}, {k: (r.get('vs_tp1') or {}).get(k) for k in ('speedup','efficiency','common_prefix_tokens_mean')})
"
