"""Ragged batches (extension): prompts of different lengths generated in one
batch must yield, sequence by sequence, exactly what each prompt yields ALONE through the oracle -
`generate` of llama3.py (pos = L + i, slot L skipped) or the llama3_simple schedule (pos = L + i - 1).
Covers the row-streaming path (B <= 8), the tensor-core paths (B > 8: swapped-role GEMM epilogues with
per-sequence positions) and per-sequence EOS."""
import numpy as np
import pytest

from llama3_np_b200 import Llama, ModelArgs
from llama3_np_b200.synth import make_weights
from oracle import ref_llama3 as orc

pytestmark = pytest.mark.gpu


def _alone(w, args, prompt, n_new, schedule):
    a1 = ModelArgs(**{**args.__dict__, "max_batch_size": 1, "dtype": "float32"})
    o = orc.OracleLlama(w, a1)
    ids = np.asarray(prompt)[None, :]
    L = ids.shape[1]
    if schedule == "llama3":
        return np.concatenate(list(o.generate(ids, L + n_new)), axis=1)[0]
    toks, nxt = [], None
    for i in range(n_new):
        lg = o(ids, 0) if i == 0 else o(nxt, L + i - 1)
        nxt = lg[:, -1, :].argmax(-1, keepdims=True)
        toks.append(int(nxt[0, 0]))
    return np.array(toks)


@pytest.mark.parametrize("schedule", ["llama3", "simple"])
@pytest.mark.parametrize("lens", [(3, 7, 5, 9), (4, 11, 2, 9, 6, 13, 8, 3, 10, 5, 7, 12)])
def test_ragged_equals_each_prompt_alone_fp32(lens, schedule):
    args = ModelArgs(dim=288, n_layers=2, n_heads=6, n_kv_heads=3, vocab_size=777, max_seq_len=48,
                     max_batch_size=len(lens))
    w = make_weights(args, 768, seed=41)
    rng = np.random.default_rng(len(lens))
    prompts = [rng.integers(3, 777, n) for n in lens]
    m = Llama(w, args)
    got = m.generate_ragged(prompts, 14, schedule=schedule)
    m.close()
    for p, g in zip(prompts, got):
        assert np.array_equal(g, _alone(w, args, p, 14, schedule)), (len(p), schedule)


def test_ragged_eos_stops_one_sequence_only():
    args = ModelArgs(dim=128, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=200, max_seq_len=40, max_batch_size=3)
    w = make_weights(args, 320, seed=42)
    rng = np.random.default_rng(7)
    prompts = [rng.integers(3, 200, n) for n in (5, 8, 3)]
    m = Llama(w, args)
    free = m.generate_ragged(prompts, 16)
    eos = int(free[1][5])                      # make sequence 1 stop at its 6th token
    m.reset_cache()
    got = m.generate_ragged(prompts, 16, eos_id=eos)
    m.close()
    for b in range(3):
        want = free[b]
        if (want == eos).any():
            want = want[: int(np.argmax(want == eos)) + 1]
        assert np.array_equal(got[b], want)
    assert len(got[1]) <= 6


def test_ragged_bf16_agrees_with_uniform_path():
    """bf16 mode: a ragged batch whose prompts happen to have equal lengths matches `generate_all`."""
    args = ModelArgs(dim=256, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=512, max_seq_len=40,
                     max_batch_size=12, dtype="bfloat16")
    w = make_weights(args, 512, seed=43)
    ids = np.random.default_rng(43).integers(3, 512, (12, 6))
    m = Llama(w, args)
    a = m.generate_all(ids, 6 + 10)
    m.reset_cache()
    b = np.stack(m.generate_ragged(list(ids), 10))
    m.close()
    assert np.array_equal(a, b)


def test_ragged_value_errors():
    args = ModelArgs(dim=64, n_layers=1, n_heads=4, vocab_size=96, max_seq_len=16, max_batch_size=2)
    m = Llama(make_weights(args, 160, seed=1), args)
    with pytest.raises(ValueError):
        m.generate_ragged([[1, 2, 3], [4]], 15)            # longest prompt + new tokens > max_seq_len
    with pytest.raises(ValueError):
        m.generate_ragged([[1], [2], [3]], 2)              # more prompts than max_batch_size
    with pytest.raises(ValueError):
        m.generate_ragged([[1, 2], []], 2)
    m.close()
