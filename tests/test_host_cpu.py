"""Host-side pieces that need no GPU: the fast tokenizer against vectors recorded from the unmodified
reference tokenizer (oracle/gen_golden_tokenizer.py), the safetensors converter, the RoPE base opt-in."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from llama3_np_b200 import convert
from llama3_np_b200.tokenizer import Tokenizer


def test_tokenizer_matches_reference_vectors():
    tok = Tokenizer(os.path.join(GOLDEN, "tokenizer_vocab_synthetic.json"))
    cases = json.load(open(os.path.join(GOLDEN, "tokenizer_cases.json"), encoding="utf-8"))
    assert len(cases) >= 30
    for c in cases:
        ids = tok.encode(c["text"], add_bos=c["add_bos"], add_eos=c["add_eos"])
        assert ids == c["ids"], c["text"]
        assert tok.decode(ids) == c["decoded"], c["text"]
    assert tok.bos_id == 1 and tok.eos_id == 2
    assert tok.str_lookup("on") == tok.vocab.index("on")  # duplicate piece: first index


@pytest.mark.skipif(not os.path.exists("/root/reference/tokenizer.model.np"), reason="reference vocabulary not on this box")
def test_tokenizer_matches_reference_on_its_own_vocabulary():
    import sys
    sys.path.insert(0, "/root/reference")
    from tokenizer import Tokenizer as Ref
    ref, mine = Ref("/root/reference/tokenizer.model.np"), Tokenizer("/root/reference/tokenizer.model.np")
    for text in ("I have a dream", "Once upon a time", "Lily thought she was the fastest girl in town.", "héllo wörld ✓"):
        ids = ref.encode(text)
        assert mine.encode(text) == ids
        assert mine.decode(ids) == ref.decode(ids)
    assert mine.encode("I have a dream") == [1, 76, 505, 263, 12561]


def test_hf_permutation_roundtrip_and_meaning():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((4 * 8, 16)).astype(np.float32)            # 4 heads, head_dim 8
    assert np.array_equal(convert.hf_unpermute(convert.hf_permute(w, 4), 4), w)
    # row (head h, interleaved index 2j + r) moves to HF row (h, r * hd/2 + j)
    p = convert.hf_permute(w, 4)
    assert np.array_equal(p[1 * 8 + 0 * 4 + 3], w[1 * 8 + 2 * 3 + 0])
    assert np.array_equal(p[2 * 8 + 1 * 4 + 1], w[2 * 8 + 2 * 1 + 1])


def test_safetensors_to_npz(tmp_path):
    from safetensors.numpy import save_file
    rng = np.random.default_rng(1)
    D, HN, KV, HD, FD, VS = 32, 4, 2, 8, 48, 40
    ref = {"model.embed_tokens.weight": rng.standard_normal((VS, D)).astype(np.float32),
           "model.norm.weight": rng.standard_normal(D).astype(np.float32)}
    for i in range(2):
        p = f"model.layers.{i}."
        ref[p + "self_attn.q_proj.weight"] = rng.standard_normal((HN * HD, D)).astype(np.float32)
        ref[p + "self_attn.k_proj.weight"] = rng.standard_normal((KV * HD, D)).astype(np.float32)
        ref[p + "self_attn.v_proj.weight"] = rng.standard_normal((KV * HD, D)).astype(np.float32)
        ref[p + "self_attn.o_proj.weight"] = rng.standard_normal((D, HN * HD)).astype(np.float32)
        for n, shp in (("up_proj", (FD, D)), ("gate_proj", (FD, D)), ("down_proj", (D, FD))):
            ref[p + f"mlp.{n}.weight"] = rng.standard_normal(shp).astype(np.float32)
        ref[p + "input_layernorm.weight"] = rng.standard_normal(D).astype(np.float32)
        ref[p + "post_attention_layernorm.weight"] = rng.standard_normal(D).astype(np.float32)
    hf = dict(ref)                                                       # what an HF export would hold
    for i in range(2):
        p = f"model.layers.{i}.self_attn."
        hf[p + "q_proj.weight"] = convert.hf_permute(ref[p + "q_proj.weight"], HN)
        hf[p + "k_proj.weight"] = convert.hf_permute(ref[p + "k_proj.weight"], KV)
    hf["model.layers.0.self_attn.rotary_emb.inv_freq"] = np.ones(4, np.float32)
    src, dst = str(tmp_path / "m.safetensors"), str(tmp_path / "m.npz")
    save_file(hf, src)                                                   # no lm_head: tied embeddings
    shapes = convert.safetensors_to_npz(src, dst, HN, KV)
    got = np.load(dst)
    assert set(got.files) == set(ref) | {"lm_head.weight"} and shapes["lm_head.weight"] == (VS, D)
    for k, v in ref.items():
        assert got[k].dtype == np.float32 and np.array_equal(got[k], v), k
    assert np.array_equal(got["lm_head.weight"], ref["model.embed_tokens.weight"])
    # the converted file is a valid weight mapping for the loader
    from llama3_np_b200.llama3 import _expected_keys
    from llama3_np_b200 import ModelArgs
    assert set(_expected_keys(ModelArgs(dim=D, n_layers=2, n_heads=HN, n_kv_heads=KV, vocab_size=VS))) == set(got.files)


def test_rope_base_is_a_parameter_of_the_table_builder():
    from llama3_np_b200.llama3 import compute_cos_sin_cache
    c1, s1 = compute_cos_sin_cache(64, 16)
    c2, s2 = compute_cos_sin_cache(64, 16, 500000.0)
    assert c1.dtype == np.float64 and c1.shape == (16, 32)
    inv = 1.0 / (500000.0 ** (np.arange(0, 64, 2) / 64))
    np.testing.assert_allclose(c2, np.cos(np.outer(np.arange(16), inv)), rtol=0, atol=1e-15)
    assert not np.allclose(c1[5], c2[5])


def test_packed_cache_key_and_header_probe(tmp_path):
    """Host side of the packed weight cache: the key is the checkpoint file's sha256, and a file that is not a pack is
    refused from its header alone (pure file IO: runs without a GPU)."""
    import ctypes as C
    import hashlib
    from llama3_np_b200 import _cabi
    from llama3_np_b200.utils import checkpoint_digest, packed_cache_path
    p = tmp_path / "w.npz"
    blob = np.random.default_rng(0).integers(0, 255, 100_000).astype(np.uint8).tobytes()
    p.write_bytes(blob)
    d = checkpoint_digest(str(p), chunk=4096)
    assert d == hashlib.sha256(blob).hexdigest()
    a = packed_cache_path(tmp_path, d, "float32")
    assert a != packed_cache_path(tmp_path, d, "bfloat16") != packed_cache_path(tmp_path, d, "bfloat16", 1, 2)
    assert os.path.basename(a).startswith(d[:40])
    lib = _cabi.lib()
    cfg, dig = _cabi.L3Config(), C.create_string_buffer(128)
    assert lib.l3_packed_info(str(p).encode(), C.byref(cfg), dig, 128) == _cabi.L3_EINVAL
    assert b"not a packed weight cache" in lib.l3_last_error(None)
    assert lib.l3_packed_info(str(tmp_path / "missing").encode(), C.byref(cfg), dig, 128) == _cabi.L3_EINVAL
