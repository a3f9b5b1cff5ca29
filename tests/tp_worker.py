"""One tensor-parallel rank of the 2-GPU parity test (launched by torchrun from test_tp_gpu.py).
Every rank loads the SAME full weight mapping, keeps its heads / FFN columns / vocabulary rows,
and must reproduce the single-device oracle: logits within the fp32 bar, tokens identical."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llama3_np_b200  # noqa: E402,F401
from llama3_np_b200 import Llama, ModelArgs, dp  # noqa: E402
from llama3_np_b200.synth import make_weights  # noqa: E402
from oracle import ref_llama3 as orc  # noqa: E402  (the checker)


def main():
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    res = {}
    for dtype, tol in (("float32", 1e-4), ("bfloat16", 3e-2)):
        for B, L in ((1, 5), (2, 12)):
            args = ModelArgs(dim=256, n_layers=3, n_heads=8, n_kv_heads=4, vocab_size=1024, max_seq_len=64,
                             max_batch_size=2, dtype=dtype)
            w = make_weights(args, 512, seed=5)
            ids = np.random.default_rng(B).integers(3, 1024, (B, L))
            oargs = ModelArgs(**{**args.__dict__, "dtype": "float32"})
            want = orc.OracleLlama(w, oargs)(ids, 0)
            want_tok = np.concatenate(list(orc.OracleLlama(w, oargs).generate(ids, 40)), axis=1)
            uid = dp.tp_unique_id(dist)  # an NCCL unique id names ONE communicator: a fresh one per model
            m = Llama(w, args, device=local, tp_rank=rank, tp_world=world, tp_unique_id=uid)
            got = m(ids, 0)
            step = m(want_tok[:, :1], L)  # one decode step on top of the prefill (GEMV + one-shot all-reduce path)
            m.reset_cache()
            got_tok = np.concatenate(list(m.generate(ids, 40)), axis=1)
            bulk = (m.reset_cache(), m.generate_all(ids, 40))[1]
            o2 = orc.OracleLlama(w, oargs)
            o2(ids, 0)
            want_step = o2(want_tok[:, :1], L)
            m.close()
            key = f"{dtype}-B{B}"
            res[key] = {"err": float(orc.scaled_max_err(got, want)), "err_step": float(orc.scaled_max_err(step, want_step)),
                        "tok_equal": bool(np.array_equal(got_tok, want_tok)), "bulk_equal": bool(np.array_equal(bulk, got_tok)),
                        "tok_agree": float((got_tok == want_tok).mean()), "tol": tol}
    # prefill-sized messages: NCCL all-reduce, the partials travelling as bf16 in bf16 mode
    for dtype, tol in (("float32", 1e-4), ("bfloat16", 3e-2)):
        args = ModelArgs(dim=256, n_layers=2, n_heads=8, n_kv_heads=4, vocab_size=1024, max_seq_len=400,
                         max_batch_size=1, dtype=dtype)
        w = make_weights(args, 512, seed=6)
        ids = np.random.default_rng(9).integers(3, 1024, (1, 300))
        want = orc.OracleLlama(w, ModelArgs(**{**args.__dict__, "dtype": "float32"}))(ids, 0)
        uid = dp.tp_unique_id(dist)
        m = Llama(w, args, device=local, tp_rank=rank, tp_world=world, tp_unique_id=uid)
        err = float(orc.scaled_max_err(m(ids, 0), want))
        m.close()
        res[f"{dtype}-long-prefill"] = {"err": err, "err_step": err, "tok_equal": True, "bulk_equal": True, "tok_agree": 1.0, "tol": tol}
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        print("TP_RESULT " + json.dumps(allres))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
