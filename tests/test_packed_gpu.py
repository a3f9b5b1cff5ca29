"""Packed device-layout weight cache: a cached load must be bit-identical to a fresh pack,
and a pack that does not belong to the checkpoint / dtype / placement must never be served."""
import ctypes as C
import os

import numpy as np
import pytest

from llama3_np_b200 import Llama, ModelArgs, _cabi
from llama3_np_b200.synth import make_weights
from llama3_np_b200.utils import checkpoint_digest, packed_cache_path

pytestmark = pytest.mark.gpu


def _checkpoint(tmp_path, seed=5, name="model.npz"):
    args = ModelArgs(dim=256, n_layers=2, n_heads=8, n_kv_heads=4, vocab_size=1024, max_seq_len=64, max_batch_size=4)
    w = make_weights(args, 512, seed=seed)
    path = os.path.join(tmp_path, name)
    np.savez(path, **w)
    return args, path


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_cached_load_is_bit_identical_to_fresh_pack(tmp_path, dtype):
    args, path = _checkpoint(str(tmp_path))
    args.dtype = dtype
    cache = os.path.join(str(tmp_path), "cache")
    ids = np.random.default_rng(3).integers(3, 1024, (4, 9))
    fresh = Llama(path, args, cache_dir=cache)
    assert not fresh.loaded_from_pack
    pack = packed_cache_path(cache, checkpoint_digest(path), dtype)
    assert os.path.exists(pack)
    want_logits = fresh.forward_f32(ids, 0)
    fresh.reset_cache()
    want_tokens = fresh.generate_all(ids, 40)
    fresh.close()

    cached = Llama(path, args, cache_dir=cache)
    assert cached.loaded_from_pack
    # (a) the device buffers: packing the cached model again gives the same file, byte for byte
    again = os.path.join(str(tmp_path), "again.l3pack")
    cached.save_packed(again, checkpoint_digest(path))
    assert open(again, "rb").read() == open(pack, "rb").read()
    # (b) what the model computes from them
    got_logits = cached.forward_f32(ids, 0)
    cached.reset_cache()
    got_tokens = cached.generate_all(ids, 40)
    cached.close()
    assert np.array_equal(got_logits.view(np.uint32), want_logits.view(np.uint32))
    assert np.array_equal(got_tokens, want_tokens)


def test_pack_of_another_checkpoint_dtype_or_placement_is_refused(tmp_path):
    args, path = _checkpoint(str(tmp_path))
    cache = os.path.join(str(tmp_path), "cache")
    Llama(path, args, cache_dir=cache).close()
    pack = packed_cache_path(cache, checkpoint_digest(path), "float32")
    lib = _cabi.lib()

    def load_into(a, digest, tp=(0, 1)):
        cfg = _cabi.L3Config(dim=a.dim, n_layers=a.n_layers, n_heads=a.n_heads, n_kv_heads=a.n_kv_heads,
                             vocab_size=a.vocab_size, max_seq_len=a.max_seq_len, max_batch_size=a.max_batch_size,
                             hidden_dim=512, norm_eps=a.norm_eps, dtype=_cabi.DTYPE_BF16 if a.dtype == "bfloat16" else _cabi.DTYPE_F32,
                             device=0, tp_rank=tp[0], tp_world=tp[1], flags=0)
        h = C.c_void_p()
        _cabi.check(lib.l3_create(C.byref(cfg), C.byref(h)))
        try:
            rc = lib.l3_load_packed(h, pack.encode(), digest.encode() if digest else None)
            msg = (lib.l3_last_error(h) or b"").decode()
            fin = lib.l3_finalize(h) if rc != 0 else None   # a refused pack leaves the model unloaded
            return rc, msg, fin
        finally:
            lib.l3_destroy(h)

    good = checkpoint_digest(path)
    assert load_into(args, good)[0] == _cabi.L3_OK
    rc, msg, fin = load_into(args, "0" * 64)
    assert rc == _cabi.L3_EINVAL and "another checkpoint" in msg and fin == _cabi.L3_ESTATE
    bf = ModelArgs(**{**args.__dict__, "dtype": "bfloat16"})
    rc, msg, _ = load_into(bf, good)
    assert rc == _cabi.L3_EINVAL and "another shape" in msg
    wide = ModelArgs(**{**args.__dict__, "n_layers": 3})
    assert load_into(wide, good)[0] == _cabi.L3_EINVAL
    assert load_into(args, good, tp=(1, 2))[0] == _cabi.L3_EINVAL


def test_damaged_pack_is_detected_and_rewritten(tmp_path):
    args, path = _checkpoint(str(tmp_path))
    cache = os.path.join(str(tmp_path), "cache")
    ids = np.random.default_rng(4).integers(3, 1024, (2, 7))
    m = Llama(path, args, cache_dir=cache)
    want = m.forward_f32(ids, 0)
    m.close()
    pack = packed_cache_path(cache, checkpoint_digest(path), "float32")
    good = open(pack, "rb").read()
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x40  # one flipped bit somewhere in the tensor data
    open(pack, "wb").write(bytes(bad))
    m = Llama(path, args, cache_dir=cache)
    assert not m.loaded_from_pack            # checksum mismatch: served from the checkpoint instead ...
    got = m.forward_f32(ids, 0)
    m.close()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert open(pack, "rb").read() == good   # ... and the pack was rewritten
    # a changed checkpoint has another digest, hence another pack: the old one is never consulted
    _, path2 = _checkpoint(str(tmp_path), seed=6, name="model2.npz")
    m = Llama(path2, args, cache_dir=cache)
    assert not m.loaded_from_pack
    m.close()
    assert len([f for f in os.listdir(cache) if f.endswith(".l3pack")]) == 2
