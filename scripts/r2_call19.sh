#!/bin/bash
# 8 GPUs: bench_tp at TP 8 (two-level sums, two-phase all-reduce), then the driver-style bench.py --gpus 8 with its tp records
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 scripts/bench_tp.py --batches 1,32 --decode 128 > gpurun_out/r2c19_tp8.log 2>&1; echo "bench_tp tp8 rc=$?"
grep '^{' gpurun_out/r2c19_tp8.log | tee gpurun_out/r2c19_tp8.jsonl | cut -c1-330
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2c19_bench_n8.log 2> gpurun_out/r2c19_bench_n8.err ) 2>&1 | grep real; echo "bench n8 done"
tail -1 gpurun_out/r2c19_bench_n8.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])
for r in d.get('tp', []):
    print({k: r.get(k) for k in ('name','tp','value','ms_per_step','ranks_agree_on_tokens','vs_tp1','error')})
"; tail -3 gpurun_out/r2c19_bench_n8.err | cut -c1-300
