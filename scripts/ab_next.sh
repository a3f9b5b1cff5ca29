#!/bin/bash
# First GPU call of the next round: measure what was written after this round's GPU budget ran out.
# Before the call, HERE (nvcc):
#   python llama3.np_b200/build.py --variant next -DL3_TC_ARGMAX_REDUX -DL3_TC_KSPLIT_PIPELINED_SUM -DL3_TC_KSPLIT_LIGHT_FENCE
#   python llama3.np_b200/build.py --variant attnu5 -DL3_ATTN_WARP_U=5
#   python llama3.np_b200/build.py --variant fuse -DL3_TC_ARGMAX_REDUX -DL3_TC_KSPLIT_PIPELINED_SUM -DL3_TC_KSPLIT_LIGHT_FENCE -DL3_TC_FUSE_NORM
# Then:  gpurun --timeout 900 -- 'bash scripts/ab_next.sh'      (about 5 GPU-minutes)
mkdir -p gpurun_out
O=gpurun_out/ab
# 1. tensor-pipe cost table (design input)
timeout 120 python scripts/mma_cost.py > ${O}_mma_cost.jsonl 2>&1; echo "mma_cost rc=$?"; tail -3 ${O}_mma_cost.jsonl
# 2. parity of the variant library (redux argmax in the LM head, prefetching K-split sum)
L3_LIB_VARIANT=next timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_mega_gpu.py tests/test_simple_gpu.py \
  tests/test_ragged_gpu.py -m gpu -q -x --timeout 150 > ${O}_pytest_next.log 2>&1; echo "pytest(next) rc=$?"; tail -2 ${O}_pytest_next.log
L3_LIB_VARIANT=fuse timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_mega_gpu.py tests/test_simple_gpu.py \
  tests/test_ragged_gpu.py -m gpu -q -x --timeout 150 > ${O}_pytest_fuse.log 2>&1; echo "pytest(fuse) rc=$?"; tail -2 ${O}_pytest_fuse.log
# 3. headline, default vs variant (same box, back to back, twice each to see the noise)
for i in 1 2; do
  timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_default_$i.log 2>&1
  L3_LIB_VARIANT=next timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_next_$i.log 2>&1
  L3_LIB_VARIANT=next L3_LM_2ACC=1 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_next_lm2acc_$i.log 2>&1
  L3_LIB_VARIANT=fuse timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_fuse_$i.log 2>&1
  L3_LIB_VARIANT=attnu5 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_attnu5_$i.log 2>&1
  L3_LIB_VARIANT=next L3_PDL=1 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > ${O}_bench_next_pdl_$i.log 2>&1
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_bench_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "tok/s  e2e", round(d["e2e"]["value"]), " step ms",
              {k.split(" ")[0]: round(v, 4) for k, v in d["roofline"]["per_decode_step_ms"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
# 4. 8B batch-32 decode: K-slices against wave quantisation of the gate|up projection
timeout 200 python scripts/bench_shapes.py 8b-b32 > ${O}_8b_b32_default.log 2>&1; python scripts/show_shapes.py ${O}_8b_b32_default.log | tail -1
L3_SWAP_TAILSPLIT=1 timeout 200 python scripts/bench_shapes.py 8b-b32 > ${O}_8b_b32_tailsplit.log 2>&1; python scripts/show_shapes.py ${O}_8b_b32_tailsplit.log | tail -1
