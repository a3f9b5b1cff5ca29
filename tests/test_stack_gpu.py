"""The cluster-resident batched-decode kernel (decode_stack.cu: every layer of a decode step inside thread-block
clusters, one cluster per block of 12 sequences) against the oracle and against the kernel-per-projection path:
identical greedy tokens in fp32 mode at the stories15M shape, including the headline configuration
(B = 256, 6 layers, vocabulary 32000), batches that leave the last cluster partly empty, contexts that need
every attention unit (positions up to max_seq_len - 1), a second generate on the same instance (stale cache
rows, llama3.py:138-153 never re-zeroes) and the llama3_simple position schedule."""
import numpy as np
import pytest

from llama3_np_b200 import Llama, ModelArgs, _cabi
from llama3_np_b200.synth import make_weights
from oracle import ref_llama3 as orc

pytestmark = pytest.mark.gpu


def _model(n_layers, vocab, max_seq, B, seed=12):
    args = ModelArgs(dim=288, n_layers=n_layers, n_heads=6, n_kv_heads=6, vocab_size=vocab, max_seq_len=max_seq,
                     max_batch_size=B)
    return args, make_weights(args, 768, seed=seed)


@pytest.mark.parametrize("B", [33, 37, 160])
def test_stack_decode_token_identical_fp32(B):
    args, w = _model(3, 2000, 40, B)
    ids = np.random.default_rng(6).integers(3, 2000, (B, 5))
    want = np.concatenate(list(orc.OracleLlama(w, args).generate(ids, 24)), axis=1)
    m = Llama(w, args)
    got = m.generate_all(ids, 24)
    assert np.array_equal(got, want)
    # one prefill (kernel per projection) + 18 decode steps of 3 launches: stack kernel, LM head, finalize
    assert m.launch_count() < 18 * 3 + 60
    m.reset_cache()
    lazy = np.concatenate(list(m.generate(ids, 24)), axis=1)
    assert np.array_equal(lazy, want)
    k_a, v_a = m.read_cache(2)
    m.close()
    m2 = Llama(w, args, flags=_cabi.FLAG_NO_MEGA)   # kernel-per-projection path
    assert np.array_equal(m2.generate_all(ids, 24), want)
    k_b, v_b = m2.read_cache(2)
    m2.close()
    np.testing.assert_allclose(k_a, k_b, rtol=0, atol=2e-5)
    np.testing.assert_allclose(v_a, v_b, rtol=0, atol=2e-5)


def test_stack_headline_config_token_identical_fp32():
    """BASELINE.json configs[1] as bench.py runs it: stories15M, 256 prompts of BOS + 7 ids, fp32 - the first
    18 generated tokens of every prompt equal the oracle's."""
    args, w = _model(6, 32000, 256, 256, seed=0)
    ids = np.random.default_rng(1).integers(3, 32000, (256, 8))
    ids[:, 0] = 1
    want = np.concatenate(list(orc.OracleLlama(w, args).generate(ids, 26)), axis=1)
    m = Llama(w, args)
    got = m.generate_all(ids, 26)
    m.close()
    assert got.shape == (256, 18)
    assert np.array_equal(got, want)


def test_stack_long_context_and_second_generate():
    """Positions up to max_seq_len - 1 = 287 (six attention units per sequence, the last one partial), then a
    second, shorter generate on the same instance: its skipped slot L still holds the first run's row."""
    args, w = _model(2, 512, 288, 36)
    o = orc.OracleLlama(w, args)
    ids = np.random.default_rng(8).integers(3, 512, (36, 3))
    want1 = np.concatenate(list(o.generate(ids, 288)), axis=1)
    ids2 = np.random.default_rng(9).integers(3, 512, (36, 9))
    want2 = np.concatenate(list(o.generate(ids2, 40)), axis=1)
    m = Llama(w, args)
    got1 = m.generate_all(ids, 288)
    got2 = m.generate_all(ids2, 40)
    m.close()
    assert np.array_equal(got1, want1)
    assert np.array_equal(got2, want2)


def test_stack_functional_surface_position_schedule():
    """llama_generate's schedule (pos = L + i - 1, llama3_simple.py:279) through the same kernel: equal to the
    kernel-per-projection path, which tests/test_simple_gpu.py pins to the reference's llama3_simple.py."""
    from llama3_np_b200 import llama3_simple as ls
    args, w = _model(2, 512, 48, 40)
    ids = np.random.default_rng(10).integers(3, 512, (40, 6))
    a = ls.llama_init(w, args)
    got = np.concatenate(list(ls.llama_generate(a, ids, 24)), axis=1)
    ls.llama_close(a)
    b = ls.llama_init(w, args, flags=_cabi.FLAG_NO_MEGA)
    ref = np.concatenate(list(ls.llama_generate(b, ids, 24)), axis=1)
    ls.llama_close(b)
    assert got.shape == (40, 24)
    assert np.array_equal(got, ref)


def test_stack_is_bitwise_deterministic():
    """Every reduction in the kernel has a fixed order, so repeated runs must agree bit for bit.  (This is the
    detector of the cross-proxy race described in profiles/r02_stack_race.txt: without the proxy fence ~30 % of
    such runs deviated.)"""
    import hashlib
    args, w = _model(6, 2000, 40, 256)
    ids = np.random.default_rng(6).integers(3, 2000, (256, 5))
    m = Llama(w, args)
    seen = set()
    for _ in range(40):
        m.reset_cache()
        tok = m.generate_all(ids, 24)
        k, v = m.read_cache(5)
        seen.add(hashlib.sha1(tok.tobytes() + k.tobytes() + v.tobytes()).hexdigest())
    m.close()
    assert len(seen) == 1
