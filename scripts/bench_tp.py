#!/usr/bin/env python
"""Tensor-parallel decode / prefill of the Llama-3-8B-shaped config (BASELINE.json configs[3]):
run under torchrun, one rank per GPU.  Heads / FFN columns / vocabulary rows are sharded, the
two row-parallel projections per layer are summed over ranks (one-shot peer-memory all-reduce
for decode-sized messages, NCCL otherwise).  Timing: CUDA events on each rank's stream, max
over ranks.  Prints one JSON line per (B) on rank 0.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      scripts/bench_tp.py [--layers 32] [--batches 1,32] [--decode 128]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200 import Llama, _cabi, dp  # noqa: E402
from llama3_np_b200.config import named_config  # noqa: E402
from llama3_np_b200.synth import param_count  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batches", default="1,32")
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--decode", type=int, default=128)
    ap.add_argument("--shape", default="llama3-8b")
    a = ap.parse_args()
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6545.0
    for B in [int(b) for b in a.batches.split(",")]:
        L, nd = a.prompt, a.decode
        args, hidden = named_config(a.shape, max_batch_size=B, max_seq_len=L + nd + 2, dtype="bfloat16", n_layers=a.layers)
        uid = dp.tp_unique_id(dist) if world > 1 else None  # one NCCL unique id per communicator
        m = Llama(None, args, hidden_dim=hidden, random_seed=0, device=local, tp_rank=rank, tp_world=world, tp_unique_id=uid)
        lib, h = m._lib, m._h
        ids = np.random.default_rng(2).integers(3, args.vocab_size, (B, L)).astype(np.int32)
        d_ids, d_out = C.c_void_p(), C.c_void_p()
        _cabi.check(lib.l3_dev_alloc(h, ids.nbytes, C.byref(d_ids)), h)
        _cabi.check(lib.l3_dev_alloc(h, B * (nd + 1) * 8, C.byref(d_out)), h)
        _cabi.check(lib.l3_memcpy_h2d(h, d_ids, ids.ctypes.data_as(C.c_void_p), ids.nbytes), h)

        def timed(fn, iters=2):
            fn()
            m.sync()
            dist.barrier()
            ms = C.c_float()
            _cabi.check(lib.l3_timer_start(h), h)
            for _ in range(iters):
                fn()
            _cabi.check(lib.l3_timer_stop(h, C.byref(ms)), h)
            return dp.max_over_ranks(ms.value / iters, dist)

        ms_pf = timed(lambda: _cabi.check(lib.l3_forward_dev(h, d_ids, B, L, 0, None, d_out), h))
        ms_gen = timed(lambda: _cabi.check(lib.l3_generate_greedy_dev(h, d_ids, B, L, L + nd, d_out), h))
        toks = np.empty((B, nd), np.int64)
        _cabi.check(lib.l3_memcpy_d2h(h, toks.ctypes.data_as(C.c_void_p), d_out, toks.nbytes), h)
        every = [None] * world
        dist.all_gather_object(every, toks[:, :8].tolist())
        ms_dec = (ms_gen - ms_pf) / (nd - 1)
        nkv, hd = args.n_kv_heads, args.dim // args.n_heads
        params = param_count(args, hidden) - args.vocab_size * args.dim + args.dim
        bytes_step = params * 2 + B * args.n_layers * nkv * (L + nd // 2) * hd * 2 * 2
        if rank == 0:
            print(json.dumps({"config": f"{a.shape}-tp{world}-b{B}", "n_layers": a.layers, "tp": world, "B": B, "prompt": L,
                              "decode": nd, "prefill_ms": ms_pf, "prefill_tok_s": B * L / ms_pf * 1e3,
                              "decode_ms_per_step": ms_dec, "decode_tok_s": B / ms_dec * 1e3,
                              "bytes_per_step_all_ranks": bytes_step,
                              "decode_hbm_frac_of_aggregate": bytes_step / (ms_dec / 1e3) / 1e9 / (hbm * world),
                              "ranks_agree": all(e == every[0] for e in every)}), flush=True)
        lib.l3_dev_free(h, d_ids)
        lib.l3_dev_free(h, d_out)
        m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
