"""GPU parity suite: the CUDA path, called through the C-ABI / the `Llama` drop-in, against
the oracle on the same seeded inputs and against the golden fixtures recorded from the
unmodified reference.

Bars (BASELINE.json north_star): fp32 mode - identical greedy tokens, logits within 1e-4 of
max|logit| ("relative" as defined in SURVEY 8(c)); bf16 mode - scaled error <= 3e-2 and
top-1 agreement on the tested steps where the oracle's top-1/top-2 gap exceeds the error.
"""
import numpy as np
import pytest

from conftest import MODEL_CASES, golden_model, load_golden
from llama3_np_b200 import Llama, ModelArgs, _cabi
from llama3_np_b200.synth import make_weights
from oracle import ref_llama3 as orc

pytestmark = pytest.mark.gpu

F32_TOL = 1e-4


def _args(args, **kw):
    d = dict(args.__dict__)
    d.update(kw)
    return ModelArgs(**d)


# ------------------------------------------------------------------------------- per-op
def test_op_rmsnorm():
    g = load_golden("ops")
    x = np.ascontiguousarray(g["rms_in"].reshape(-1, 288))
    out = np.empty_like(x)
    assert _cabi.lib().l3_op_rmsnorm(0, _cabi.f32p(x), _cabi.f32p(g["rms_w"]), 1e-6, x.shape[0], 288, _cabi.f32p(out)) == 0
    np.testing.assert_allclose(out, g["rms_out"].reshape(-1, 288), rtol=2e-6, atol=1e-6)


def test_op_rope_matches_reference_outputs():
    g = load_golden("ops")
    for x, want in ((g["rope_q_in"], g["rope_q_out"]), (g["rope_k_in"], g["rope_k_out"])):
        B, L, H, HD = x.shape
        out = np.empty_like(x)
        rc = _cabi.lib().l3_op_rope(0, _cabi.f32p(np.ascontiguousarray(x)), _cabi.f64p(np.ascontiguousarray(g["cos48"])),
                                    _cabi.f64p(np.ascontiguousarray(g["sin48"])), B, L, H, HD, 4, _cabi.f32p(out))
        assert rc == 0
        np.testing.assert_allclose(out, want, rtol=0, atol=2e-6)


def test_op_swiglu():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(5000).astype(np.float32) * 3
    b = rng.standard_normal(5000).astype(np.float32)
    out = np.empty_like(a)
    assert _cabi.lib().l3_op_swiglu(0, _cabi.f32p(a), _cabi.f32p(b), a.size, _cabi.f32p(out)) == 0
    want = orc.silu(a.astype(np.float64)) * b
    np.testing.assert_allclose(out, want, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("rows,n,k,path", [
    (1, 288, 288, 1), (1, 32000, 288, 1), (1, 1536, 288, 1), (1, 288, 768, 1), (3, 301, 288, 1),
    (8, 512, 1024, 1), (1, 4096, 4096, 1), (2, 1024, 14336, 1),
    (5, 96, 64, 2), (64, 288, 288, 2), (256, 1536, 288, 2), (130, 333, 768, 2), (256, 32000, 288, 2),
])
def test_op_linear_fp32(rows, n, k, path):
    rng = np.random.default_rng(rows * 7 + n)
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    assert _cabi.lib().l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, path, 0, _cabi.f32p(out)) == 0
    want = x.astype(np.float64) @ w.astype(np.float64).T
    assert orc.scaled_max_err(out, want) < 2e-6


@pytest.mark.parametrize("rows,n,k", [
    (128, 64, 64), (256, 288, 288), (256, 1536, 288), (256, 288, 768), (200, 333, 96), (9, 40, 32),
    (256, 32000, 288), (300, 1024, 2048), (2048, 512, 1024),
    (1024, 1280, 1024),   # more output tiles than SMs: persistent CTAs walk a second, partly filled wave
])
def test_op_linear_tcgen05_tf32x3(rows, n, k):
    """fp32-mode tensor-core GEMM: 3xTF32 split must stay at fp32-level accuracy."""
    rng = np.random.default_rng(rows + n + k)
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    assert _cabi.lib().l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, 3, 0, _cabi.f32p(out)) == 0
    want = x.astype(np.float64) @ w.astype(np.float64).T
    assert orc.scaled_max_err(out, want) < 5e-6


@pytest.mark.parametrize("rows,n,k", [(128, 64, 64), (256, 288, 288), (130, 1000, 776), (2048, 768, 1024), (40, 256, 4096),
                                      # more output tiles than SMs (several waves per persistent CTA, ragged edges)
                                      (2048, 3072, 1024), (1024, 6144, 2048), (1000, 5000, 1544)])
def test_op_linear_tcgen05_bf16(rows, n, k):
    import torch
    rng = np.random.default_rng(rows + n + k)
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    assert _cabi.lib().l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, 3, 1, _cabi.f32p(out)) == 0
    xb = torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()
    wb = torch.from_numpy(w).to(torch.bfloat16).to(torch.float64).numpy()
    assert orc.scaled_max_err(out, xb @ wb.T) < 5e-6  # exact products of bf16 inputs, fp32 accumulation


@pytest.mark.parametrize("rows,path", [(1, 1), (4, 1), (40, 2)])
def test_op_linear_bf16_weights(rows, path):
    import torch
    rng = np.random.default_rng(rows)
    x = rng.standard_normal((rows, 512)).astype(np.float32)
    w = (rng.standard_normal((300, 512)) / 22).astype(np.float32)
    out = np.empty((rows, 300), np.float32)
    assert _cabi.lib().l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, 300, 512, path, 1, _cabi.f32p(out)) == 0
    wb = torch.from_numpy(w).to(torch.bfloat16).to(torch.float64).numpy()  # same rounding as the device
    want = x.astype(np.float64) @ wb.T
    assert orc.scaled_max_err(out, want) < 2e-6


@pytest.mark.parametrize("B,L,HN,KVHN,HD,start,nsplit", [
    (2, 1, 6, 6, 48, 37, 1), (2, 1, 6, 6, 48, 200, 4), (3, 1, 8, 2, 64, 129, 3), (1, 1, 32, 8, 128, 300, 8),
    (2, 1, 4, 4, 16, 9, 1), (1, 1, 6, 2, 32, 70, 2), (1, 1, 6, 2, 96, 33, 1), (2, 1, 8, 1, 64, 50, 1),
    (2, 5, 6, 6, 48, 0, 1), (1, 40, 8, 2, 64, 0, 1), (2, 19, 4, 1, 128, 23, 1), (1, 100, 4, 4, 16, 7, 1),
    (160, 1, 4, 4, 48, 37, 1), (150, 1, 16, 4, 64, 90, 1), (256, 1, 6, 6, 48, 133, 1),  # one warp per (sequence, head group)
    # >= 148 CTAs: key ranges staged in shared memory by bulk copies (refilled stages, partly filled stages, empty splits)
    (32, 1, 32, 8, 128, 300, 2), (40, 1, 8, 2, 64, 129, 3), (64, 1, 6, 6, 48, 200, 4), (100, 1, 4, 2, 16, 9, 1),
    (50, 1, 8, 2, 32, 4, 4), (20, 1, 16, 2, 128, 1000, 4), (30, 1, 12, 6, 96, 75, 1),
])
def test_op_attention_fp32(B, L, HN, KVHN, HD, start, nsplit):
    rng = np.random.default_rng(B * 100 + L + HD)
    T = start + L
    q = rng.standard_normal((B, L, HN, HD)).astype(np.float32)
    k = rng.standard_normal((B, T, KVHN, HD)).astype(np.float32)
    v = rng.standard_normal((B, T, KVHN, HD)).astype(np.float32)
    out = np.empty((B, L, HN * HD), np.float32)
    rc = _cabi.lib().l3_op_attention(0, _cabi.f32p(q), _cabi.f32p(k), _cabi.f32p(v), B, L, HN, KVHN, HD, start, 0,
                                     nsplit, _cabi.f32p(out))
    assert rc == 0
    # oracle: the attention core of llama3.py:190-207 in float64
    nrep = HN // KVHN
    kk = np.repeat(k.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    vv = np.repeat(v.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    s = q.astype(np.float64).transpose(0, 2, 1, 3) @ kk.transpose(0, 1, 3, 2) / np.sqrt(HD)
    if L > 1:
        mask = np.concatenate([np.zeros((L, start)), np.triu(np.full((L, L), -np.inf), k=1)], axis=1)
        s = s + mask[None, None]
    want = (orc.softmax_lastdim(s) @ vv).transpose(0, 2, 1, 3).reshape(B, L, -1)
    assert orc.scaled_max_err(out, want) < 3e-6


@pytest.mark.parametrize("B,HN,KVHN,HD,start,nsplit", [(32, 32, 8, 128, 255, 2), (32, 32, 8, 128, 383, 2), (48, 16, 4, 64, 40, 1),
                                                       (3, 8, 2, 64, 129, 3), (9, 32, 4, 128, 700, 8), (40, 8, 2, 64, 4, 4)])
@pytest.mark.parametrize("kernel", ["exact", "model"])
def test_op_attention_decode_bf16_cache(B, HN, KVHN, HD, start, nsplit, kernel):
    """Decode attention over a bf16 cache against float64 on the bf16-rounded K / V.  `exact`: the lane-group kernels
    (staged key ranges at >= 148 CTAs, else streamed from global memory) - exact products, fp32 accumulation.  `model`:
    what the model runs - for GQA groups of 4 / 8 heads at head_dim 64 / 128 and >= 148 CTAs the mma.sync kernel, which
    rounds q and the softmax weights to bf16 (refilled stages, partly filled tiles, empty splits, 1 .. 8 splits)."""
    import torch
    rng = np.random.default_rng(B + HD + start)
    T = start + 1
    q = rng.standard_normal((B, 1, HN, HD)).astype(np.float32)
    k = torch.from_numpy(rng.standard_normal((B, T, KVHN, HD)).astype(np.float32)).to(torch.bfloat16).float().numpy()
    v = torch.from_numpy(rng.standard_normal((B, T, KVHN, HD)).astype(np.float32)).to(torch.bfloat16).float().numpy()
    out = np.empty((B, 1, HN * HD), np.float32)
    rc = _cabi.lib().l3_op_attention(0, _cabi.f32p(q), _cabi.f32p(k), _cabi.f32p(v), B, 1, HN, KVHN, HD, start,
                                     3 if kernel == "exact" else 1, nsplit, _cabi.f32p(out))
    assert rc == 0
    nrep = HN // KVHN
    kk = np.repeat(k.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    vv = np.repeat(v.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    s = q.astype(np.float64).transpose(0, 2, 1, 3) @ kk.transpose(0, 1, 3, 2) / np.sqrt(HD)
    want = (orc.softmax_lastdim(s) @ vv).transpose(0, 2, 1, 3).reshape(B, 1, -1)
    assert orc.scaled_max_err(out, want) < (3e-6 if kernel == "exact" else 1e-2)


def test_op_argmax_first_maximum_wins():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 32000)).astype(np.float32)
    x[0, 777] = x[0, 31999] = 9.0   # tie -> first index
    x[1, :] = -np.inf               # all -inf -> 0
    x[2, 0] = 50.0
    x[3, 31999] = 50.0
    out = np.empty(5, np.int64)
    assert _cabi.lib().l3_op_argmax(0, _cabi.f32p(x), 5, 32000, _cabi.i64p(out)) == 0
    assert np.array_equal(out, x.argmax(-1))


# ------------------------------------------------------------------------------- model, fp32
@pytest.mark.parametrize("name", MODEL_CASES)
def test_forward_matches_reference_fixture_fp32(name):
    args, hidden, w, g = golden_model(name)
    ids = g["ids"]
    B, L = ids.shape
    m = Llama(w, args)
    a = m(ids, 0)
    assert a.dtype == np.float64 and a.shape == (B, 1, args.vocab_size)
    assert orc.scaled_max_err(a, g["logits_prefill"]) < F32_TOL
    assert orc.scaled_max_err(m(g["nxt"], L), g["logits_decode0"]) < F32_TOL
    assert orc.scaled_max_err(m(g["nxt2"], L + 1), g["logits_decode1"]) < F32_TOL
    m2 = Llama(w, args)
    m2(ids[:, :3], 0)
    assert orc.scaled_max_err(m2(ids[:, 3:], 3), g["logits_chunked"]) < F32_TOL
    m.close(); m2.close()


@pytest.mark.parametrize("name", MODEL_CASES)
def test_generate_token_identical_fp32(name):
    args, hidden, w, g = golden_model(name)
    ids, cap = g["ids"], int(g["cap"])
    B, L = ids.shape
    m = Llama(w, args)
    gen = m.generate(ids, cap)
    first = next(gen)
    assert first.shape == (B, 1) and first.dtype == np.int64
    toks = np.concatenate([first] + list(gen), axis=1)
    assert np.array_equal(toks, g["tokens"])
    # second generate on the same instance (stale cache, shorter prompt) - bulk device loop
    toks2 = m.generate_all(ids[:, : L - 2], cap - 3)
    assert np.array_equal(toks2, g["tokens_second"])
    k, v = m.read_cache(0)
    np.testing.assert_allclose(k, g["cache_k0"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(v, g["cache_v0"], rtol=0, atol=2e-5)
    m.close()


def test_position_quirk_on_device():
    args, hidden, w, g = golden_model("tiny_gqa")
    ids = g["ids"]
    B, L = ids.shape
    m = Llama(w, args)
    m.generate_all(ids, L + 4)
    k, _ = m.read_cache(0)
    assert np.all(k[:B, L] == 0) and np.any(k[:B, L + 1] != 0) and np.any(k[:B, L - 1] != 0)
    m.close()


def test_stories15m_c1_token_identical():
    """BASELINE.json configs[0]: 'I have a dream', cap 50, batch 1 -> 45 tokens."""
    args, hidden, w, g = golden_model("stories15m_c1")
    m = Llama(w, args)
    logits = m(g["ids"], 0)
    assert orc.scaled_max_err(logits, g["logits_prefill"].astype(np.float64)) < F32_TOL
    assert np.array_equal(np.argsort(-logits[0, 0])[:5], g["top5"])
    m.reset_cache()
    toks = np.concatenate(list(m.generate(g["ids"], 50)), axis=1)
    assert np.array_equal(toks, g["tokens"])
    m.reset_cache()
    assert np.array_equal(m.generate_all(g["ids"], 50), g["tokens"])
    m.close()


def test_batched_decode_vs_live_oracle_fp32():
    """configs[1] in small: B independent equal-length prompts decode exactly as B separate
    batch-1 runs of the oracle (tiled-GEMM path: B*L > 8 rows)."""
    args = ModelArgs(dim=288, n_layers=2, n_heads=6, vocab_size=1000, max_seq_len=40, max_batch_size=24)
    w = make_weights(args, 768, seed=5)
    rng = np.random.default_rng(5)
    ids = np.concatenate([np.ones((24, 1), np.int64), rng.integers(3, 1000, (24, 7))], axis=1)
    o = orc.OracleLlama(w, args)
    want = np.concatenate(list(o.generate(ids, 40)), axis=1)
    m = Llama(w, args)
    got = m.generate_all(ids, 40)
    assert np.array_equal(got, want)
    m.reset_cache()
    got2 = np.concatenate(list(m.generate(ids, 40)), axis=1)
    assert np.array_equal(got2, want)
    m.close()


def test_simt_and_tensor_core_paths_agree_fp32():
    """L3_FLAG_NO_TENSORCORE keeps the FFMA GEMMs: both paths must match the oracle."""
    args = ModelArgs(dim=288, n_layers=2, n_heads=6, n_kv_heads=3, vocab_size=777, max_seq_len=40, max_batch_size=16)
    w = make_weights(args, 768, seed=6)
    ids = np.random.default_rng(6).integers(0, 777, (16, 9))
    want = orc.OracleLlama(w, args)(ids, 0)
    for flags in (0, _cabi.FLAG_NO_TENSORCORE, _cabi.FLAG_NO_GRAPH):
        m = Llama(w, args, flags=flags)
        assert orc.scaled_max_err(m(ids, 0), want) < F32_TOL
        m.close()


def test_chunked_long_prompt_equals_single_pass():
    """Prompts longer than the workspace are processed in chunks; logits must not change."""
    args = ModelArgs(dim=64, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=128, max_seq_len=200, max_batch_size=1)
    w = make_weights(args, 160, seed=9)
    ids = np.random.default_rng(9).integers(0, 128, (1, 150))
    want = orc.OracleLlama(w, args)(ids, 0)
    m = Llama(w, args)
    assert orc.scaled_max_err(m(ids, 0), want) < F32_TOL
    m.close()


def test_value_errors():
    args, hidden, w, g = golden_model("tiny_mha")
    m = Llama(w, args)
    with pytest.raises(ValueError):
        m(np.zeros((args.max_batch_size + 1, 2), np.int64), 0)
    with pytest.raises(ValueError):
        m(np.zeros((1, 4), np.int64), args.max_seq_len - 2)
    with pytest.raises(ValueError):
        m(np.full((1, 2), args.vocab_size), 0)
    with pytest.raises(ValueError):
        list(m.generate(np.zeros((1, 2), np.int64), args.max_seq_len + 1))
    assert list(m.generate(np.zeros((1, 5), np.int64), 5)) == []
    m.close()


# ------------------------------------------------------------------------------- model, bf16
@pytest.mark.parametrize("name", ["tiny_gqa", "hd48_gqa", "hd128_gqa"])
def test_forward_bf16_tolerance(name):
    args, hidden, w, g = golden_model(name)
    ids = g["ids"]
    B, L = ids.shape
    m = Llama(w, _args(args, dtype="bfloat16"))
    a = m(ids, 0)
    err = orc.scaled_max_err(a, g["logits_prefill"])
    assert err < 3e-2, err
    d0 = m(g["nxt"], L)
    assert orc.scaled_max_err(d0, g["logits_decode0"]) < 3e-2
    # top-1 agreement wherever the oracle's margin exceeds the bf16 error
    ref = g["logits_prefill"][:, 0]
    srt = np.sort(ref, axis=-1)
    safe = (srt[:, -1] - srt[:, -2]) > 2 * err * np.abs(ref).max()
    assert safe.any(), "no row has a top-1 margin above the bf16 error: the agreement check would be vacuous"
    assert np.array_equal(a[:, 0].argmax(-1)[safe], ref.argmax(-1)[safe])
    m.close()


# ------------------------------------------------------------------------------- tensor-core prefill attention
@pytest.mark.parametrize("B,L,HN,KVHN,HD,start", [
    (1, 128, 4, 1, 64, 0), (2, 300, 8, 2, 64, 0), (1, 257, 4, 4, 128, 0), (2, 130, 8, 2, 128, 70),
    (1, 40, 4, 1, 128, 0), (1, 1000, 2, 1, 64, 24),
])
def test_op_attention_tcgen05_prefill(B, L, HN, KVHN, HD, start):
    """bf16 tensor-core flash prefill (attention_tc.cu) against the float64 attention core of
    llama3.py:190-207 evaluated on the same bf16-rounded q, k, v."""
    import torch
    rng = np.random.default_rng(B * 1000 + L + HD)
    T = start + L
    r16 = lambda x: torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    q = r16(rng.standard_normal((B, L, HN, HD)).astype(np.float32))
    k = r16(rng.standard_normal((B, T, KVHN, HD)).astype(np.float32))
    v = r16(rng.standard_normal((B, T, KVHN, HD)).astype(np.float32))
    out = np.empty((B, L, HN * HD), np.float32)
    rc = _cabi.lib().l3_op_attention(0, _cabi.f32p(q), _cabi.f32p(k), _cabi.f32p(v), B, L, HN, KVHN, HD, start, 2, 0,
                                     _cabi.f32p(out))
    assert rc == 0
    nrep = HN // KVHN
    kk = np.repeat(k.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    vv = np.repeat(v.astype(np.float64), nrep, axis=2).transpose(0, 2, 1, 3)
    s = q.astype(np.float64).transpose(0, 2, 1, 3) @ kk.transpose(0, 1, 3, 2) / np.sqrt(HD)
    mask = np.concatenate([np.zeros((L, start)), np.triu(np.full((L, L), -np.inf), k=1)], axis=1)
    want = (orc.softmax_lastdim(s + mask[None, None]) @ vv).transpose(0, 2, 1, 3).reshape(B, L, -1)
    # P and the output are rounded to bf16 (2^-9 relative each)
    assert orc.scaled_max_err(out, want) < 1.5e-2


@pytest.mark.parametrize("hd,heads,kv", [(64, 4, 2), (128, 2, 1)])
def test_forward_bf16_long_prompt_tensor_core_attention(hd, heads, kv):
    """bf16 model forward with a multi-block prompt (tcgen05 GEMMs + tcgen05 flash attention),
    single pass and chunked, against the oracle."""
    dim = hd * heads
    args = ModelArgs(dim=dim, n_layers=2, n_heads=heads, n_kv_heads=kv, vocab_size=512, max_seq_len=400,
                     max_batch_size=2, dtype="bfloat16")
    w = make_weights(args, 2 * dim, seed=13)
    ids = np.random.default_rng(13).integers(0, 512, (2, 300))
    want = orc.OracleLlama(w, _args(args, dtype="float32"))(ids, 0)
    m = Llama(w, args)
    assert orc.scaled_max_err(m(ids, 0), want) < 3e-2
    m.reset_cache()
    m(ids[:, :170], 0)
    assert orc.scaled_max_err(m(ids[:, 170:], 170), want) < 3e-2
    m.close()


# ------------------------------------------------------------------------------- swapped-role GEMM (9..128 rows)
@pytest.mark.parametrize("rows,n,k,bf16", [
    (32, 4096, 4096, 1), (9, 256, 512, 1), (40, 1000, 776, 1), (128, 1536, 288, 1), (100, 333, 2048, 1),
    (24, 288, 288, 0), (32, 864, 288, 0), (70, 32000, 288, 0), (128, 768, 1024, 0), (17, 130, 96, 0),
    # a ragged second wave (150 / 149 row blocks), the 8B QKV and Wdown shapes (48 / 32 row blocks: K-split)
    (32, 19200, 1024, 1), (24, 19072, 288, 0), (64, 6144, 4096, 1), (128, 4096, 14336, 1), (12, 2048, 2048, 0),
])
def test_op_linear_tcgen05_swapped_roles(rows, n, k, bf16):
    """gemm_swap.cu: weights as the 128-row MMA operand, the batch as N; fp32 mode via 3xTF32.  Matrices with few row
    blocks are split along K, the slices of a row block summed in slice order."""
    import torch
    rng = np.random.default_rng(rows + n + k)
    x = rng.standard_normal((rows, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    out = np.empty((rows, n), np.float32)
    assert _cabi.lib().l3_op_linear(0, _cabi.f32p(x), _cabi.f32p(w), rows, n, k, 4, bf16, _cabi.f32p(out)) == 0
    if bf16:
        x = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
        w = torch.from_numpy(w).to(torch.bfloat16).to(torch.float32).numpy()
    want = x.astype(np.float64) @ w.astype(np.float64).T
    assert orc.scaled_max_err(out, want) < 5e-6


def test_batched_decode_bf16_tolerance_swapped_gemm():
    """B = 24 decode steps in bf16 mode run the swapped-role GEMMs (incl. the K-split residual
    projections): logits of a decode step stay within the bf16 bar of the oracle."""
    args = ModelArgs(dim=512, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=1024, max_seq_len=48,
                     max_batch_size=24, dtype="bfloat16")
    w = make_weights(args, 4608, seed=17)
    ids = np.random.default_rng(17).integers(3, 1024, (24, 6))
    o = orc.OracleLlama(w, _args(args, dtype="float32"))
    o(ids, 0)
    nxt = np.random.default_rng(18).integers(3, 1024, (24, 1))
    want = o(nxt, 6)
    m = Llama(w, args)
    m(ids, 0)
    got = m(nxt, 6)
    assert orc.scaled_max_err(got, want) < 3e-2
    m.close()


def test_batched_decode_ksplit_swapped_gemm_token_identical_fp32():
    """B = 16 decode at an 8B-like aspect ratio (few weight row blocks, long K): every swapped-role GEMM of
    the step - QKV with RoPE, Wo / Wdown with the residual, gate|up with SwiGLU, the LM head with the fused
    argmax - runs K-split with the deterministic slice-order reduction; tokens must equal the oracle's."""
    args = ModelArgs(dim=1024, n_layers=2, n_heads=8, n_kv_heads=2, vocab_size=1024, max_seq_len=40, max_batch_size=16)
    w = make_weights(args, 4608, seed=19)
    ids = np.random.default_rng(19).integers(3, 1024, (16, 6))
    want = np.concatenate(list(orc.OracleLlama(w, args).generate(ids, 30)), axis=1)
    m = Llama(w, args)
    got = m.generate_all(ids, 30)
    assert np.array_equal(got, want)
    m.reset_cache()
    assert orc.scaled_max_err(m(ids, 0), orc.OracleLlama(w, args)(ids, 0)) < F32_TOL
    m.close()


def test_forward_between_generate_steps_leaves_the_loop_alone():
    """A forward call (with argmax) between l3_generate_begin and the first step, and between two steps, must not change
    the generated stream: the forward entry points own their id / argmax buffers (advisor finding, round 1)."""
    args = ModelArgs(dim=64, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=300, max_seq_len=48, max_batch_size=2)
    w = make_weights(args, 128, seed=31)
    ids = np.array([[5, 6, 7, 8], [9, 10, 11, 12]])
    other = np.array([[200, 201, 202, 203, 204, 205], [17, 18, 19, 20, 21, 22]])
    m = Llama(w, args)
    want = np.concatenate(list(m.generate(ids, 20)), axis=1)
    probe = Llama(w, args)   # the forward calls go to positions the generate never reads back (start_pos 30)
    want_probe = probe.forward_f32(other, 30, want_argmax=True)[1]
    probe.close()
    m.reset_cache()
    gen = m.generate(ids, 20)
    got = []
    _cabi.check(m._lib.l3_generate_begin(m._h, _cabi.i32p(np.ascontiguousarray(ids, dtype=np.int32)), 2, 4), m._h)
    assert np.array_equal(m.forward_f32(other, 30, want_argmax=True)[1], want_probe)   # prompt still pending in d_ids
    for i in range(16):
        nxt = np.empty((2,), np.int64)
        _cabi.check(m._lib.l3_generate_next(m._h, _cabi.i64p(nxt)), m._h)
        got.append(nxt[:, None])
        if i in (0, 5):
            m.forward_f32(other, 30, want_argmax=True)
    del gen
    m.close()
    assert np.array_equal(np.concatenate(got, axis=1), want)


def test_forward_long_prompt_more_tiles_than_sms():
    """Prefill whose projections have more output tiles than the GPU has SMs (QKV 224, Wo / W2 160, gate|up 640 tiles of
    128 x 256 at 4096 rows): every persistent CTA walks several tiles through both TMEM accumulator buffers behind the
    RoPE + cache, residual and SwiGLU epilogues; bf16 and fp32 (3xTF32) modes against the oracle."""
    args = ModelArgs(dim=1280, n_layers=2, n_heads=10, n_kv_heads=2, vocab_size=512, max_seq_len=2064, max_batch_size=2)
    w = make_weights(args, 2560, seed=41)
    ids = np.random.default_rng(41).integers(0, 512, (2, 2048))
    o = orc.OracleLlama(w, args)
    want = o(ids, 0)
    want_k = o.layers[1]["cache_k"][:2, :2048].astype(np.float32)
    m = Llama(w, args)
    got = m(ids, 0)
    k1, _ = m.read_cache(1)
    m.close()
    assert orc.scaled_max_err(got, want) < F32_TOL
    assert orc.scaled_max_err(k1[:, :2048], want_k) < F32_TOL
    mb = Llama(w, _args(args, dtype="bfloat16"))
    err = orc.scaled_max_err(mb(ids, 0), want)
    mb.close()
    assert err < 3e-2, err
