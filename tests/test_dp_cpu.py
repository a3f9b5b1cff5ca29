"""World-size-2 gloo test of the data-parallel host logic (no GPU): shards partition the
prompts, the timing reduction takes the max, gathered tokens come back in prompt order."""
import os
import socket

import numpy as np
import pytest

from llama3_np_b200 import dp


def test_shard_rows_partition():
    for n in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            blocks = [dp.shard_rows(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = np.arange(10 * 3).reshape(10, 3)
    mine = dp.shard_prompts(ids, rank, world)
    toks = mine * 2 + 1                      # stand-in for each rank's generate output
    slow = dp.max_over_ranks(10.0 + rank, dist)
    allt = dp.gather_tokens(toks, dist)
    dist.barrier()
    if rank == 0:
        q.put((slow, allt))
    dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    slow, allt = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slow == 11.0
    assert np.array_equal(allt, np.arange(30).reshape(10, 3) * 2 + 1)


def test_tp_shard_shapes_cover_the_matrices():
    # 8B shape over 8 ranks: 4 Q heads + 1 KV head per rank, FFN 1792 columns, 16032 vocabulary rows
    s = dp.tp_shard_shapes(4096, 32, 8, 14336, 128256, 8)
    assert s["wqkv"] == ((4 + 2) * 128, 4096) and s["wo"] == (4096, 512)
    assert s["w13"] == (2 * 1792, 4096) and s["w2"] == (4096, 1792)
    assert s["lm_head"] == (16032, 4096) and s["kv_cache_heads"] == 1
    with pytest.raises(ValueError):
        dp.tp_shard_shapes(4096, 32, 8, 14336, 128256, 3)


def _uid_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeLib:  # rank 0 "creates" the id; no NCCL / GPU in this container
        @staticmethod
        def l3_nccl_unique_id(buf):
            buf.raw = bytes(range(128))
            return 0
    from llama3_np_b200 import _cabi
    _cabi._lib = FakeLib()
    uid = dp.tp_unique_id(dist)
    q.put((rank, uid))
    dist.barrier()
    dist.destroy_process_group()


def test_tp_unique_id_broadcast_two_ranks():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_uid_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1] == bytes(range(128))
