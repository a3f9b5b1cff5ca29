"""The functional surface (`llama3_np_b200.llama3_simple`) against outputs recorded from the
unmodified reference `llama3_simple.py` (tests/golden/simple_*.npz, oracle/gen_golden_simple.py):
same logits within the reference's own tolerance (rtol 2e-4, atol 1e-4,
tests/test_llama_implementations.py:23-24,176-179) and identical tokens under its position
schedule; plus GQA, which the reference's functional file cannot run, against the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from llama3_np_b200 import ModelArgs
from llama3_np_b200 import llama3_simple as ls
from llama3_np_b200.synth import make_weights
from oracle import ref_llama3 as orc

pytestmark = pytest.mark.gpu


def _case(name):
    g = load_golden(name)
    fields = {k[4:]: (None if int(g[k]) == -1 else int(g[k])) for k in g.files if k.startswith("cfg_")}
    args = ModelArgs(**fields)
    return g, args, make_weights(args, int(g["hidden"]), int(g["seed"]))


@pytest.mark.parametrize("name", ["simple_tiny_mha", "simple_hd48_mha"])
def test_functional_surface_matches_reference_simple(name):
    g, args, w = _case(name)
    m = ls.llama_init(w, args)
    logits = ls.llama_forward(m, g["ids"], 0)
    assert logits.dtype == np.float32 and logits.shape == g["logits_prefill"].shape
    np.testing.assert_allclose(logits, g["logits_prefill"], rtol=2e-4, atol=1e-4)
    assert np.array_equal(np.argsort(-logits[:, 0], axis=-1)[:, :5], np.argsort(-g["logits_prefill"][:, 0], axis=-1)[:, :5])
    ls.llama_close(m)
    m = ls.llama_init(w, args)
    toks = list(ls.llama_generate(m, g["ids"], int(g["max_new_tokens"])))
    assert toks[0].dtype == np.int32 and toks[0].shape == (g["ids"].shape[0], 1)
    assert np.array_equal(np.concatenate(toks, axis=1), g["tokens"])   # stops at max_seq_len like the reference
    k, _ = m["_llama"].read_cache(0)
    L = g["ids"].shape[1]
    assert np.any(k[: g["ids"].shape[0], L] != 0)                       # no skipped slot in this schedule
    ls.llama_close(m)


def test_functional_surface_gqa_and_float16_mode():
    args = ModelArgs(dim=256, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=512, max_seq_len=40, max_batch_size=2)
    w = make_weights(args, 512, seed=31)
    ids = np.random.default_rng(31).integers(0, 512, (2, 6))
    o = orc.OracleLlama(w, args)
    want, nxt, cur = [], None, 6
    for i in range(20):
        lg = o(ids, 0) if i == 0 else o(nxt, 6 + i - 1)
        nxt = lg[:, -1, :].argmax(-1, keepdims=True)
        want.append(nxt)
    m = ls.llama_init(w, args)
    got = np.concatenate(list(ls.llama_generate(m, ids, 20)), axis=1)
    assert np.array_equal(got, np.concatenate(want, axis=1))
    ls.llama_close(m)
    args16 = ModelArgs(**{**args.__dict__, "dtype": "float16"})
    m = ls.llama_init(w, args16)
    lg = ls.llama_forward(m, ids, 0)
    assert lg.dtype == np.float16
    assert orc.scaled_max_err(lg.astype(np.float64), orc.OracleLlama(w, args)(ids, 0)) < 3e-2
    ls.llama_close(m)
