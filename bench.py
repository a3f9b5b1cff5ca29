#!/usr/bin/env python
"""bench.py - BASELINE.json's headline metric on its configs[1], plus every other north-star number as sub-records.

Headline (the JSON line's top level):
  stories15M (random-init weights in the reference's .npz layout), batched greedy decode of 256 independent
  prompts per GPU (BOS + 7 ids each, `generate(ids, 256)` -> 248 tokens per prompt), prompts sharded
  data-parallel over the GPUs with no data-path collective.  A "step" is one full `generate` over the rank's
  batch (one prefill + 247 decode steps).  `value` = generated tokens/s of the whole job with the prompt ids
  already resident in HBM (l3_generate_greedy_dev); `e2e` = the same through the drop-in `Llama.generate`
  generator with host ids (H2D inside) and every yielded token read back to the host.

`extra` (N = 1 only; each with its own clock record, roofline fraction of the MEASURED peak, an end-to-end figure
through `Llama.generate` / `Llama.__call__`, and a CPU baseline of the oracle port timed beside it):
  s15m_b1        configs[0]: stories15M, 'I have a dream', batch 1, total 50 and 256 tokens
  1b             configs[2]: Llama-3.2-1B-shaped bf16, prefill 2048 + 256 decode
  8b_b1, 8b_b32  configs[3] on one GPU: Llama-3-8B-shaped bf16, 128-token prompts + 256 decode
  8b_prefill     north_star: 8B-shaped bf16 prefill of 2048 tokens (tensor roofline)
`tp` (N > 1): configs[3] / [4] tensor-parallel over the N GPUs - 8B bf16 batch-1 and batch-32 decode, and the
  32768-token prefill at N = 8 - tokens compared with the other ranks, efficiency against the ideal N x one GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|bf16] [--impl reference] [--no-extra]

Under torchrun (N > 1) each rank drives one GPU; timing is CUDA events on the library's own stream, max over
ranks (the ranks meet on a gloo group: nobody spins in NCCL while rank 0 times its CPU leg).  `--impl reference`
times the CPU oracle port of the reference (oracle/ref_llama3.py; the reference is NumPy-only Python, nothing to
compile) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT_LEN = 8
TOTAL_LEN = 256
N_OUT = TOTAL_LEN - PROMPT_LEN  # 248 yielded tokens per prompt
# library switches that change which kernels run: recorded in the JSON line so that a number names its path
ENV_SWITCHES = ("L3_LIB_VARIANT", "L3_STACK", "L3_STACK_MIN_B", "L3_MEGA", "L3_PDL", "L3_GEMM_SWAP", "L3_GEMM_KSPLIT",
                "L3_LM_2ACC", "L3_LM_BN", "L3_ATTN_TC", "L3_TP_BF16_AR", "L3_SWAP_RESID_ATOMIC", "L3_CARVEOUT",
                "L3_STACK_PF", "L3_STACK_KV_EVICT_FIRST", "L3_TP_ONESHOT", "L3_TP_TIMEOUT_MS", "L3_ATTN_MMA", "L3_ATTN_MMA_NW",
                "L3_ATTN_STAGED", "L3_ATTN_TARGET_CTAS", "L3_PDL_GEMM")


def set_total_len(n):
    """--total-len: shorten the generate for profiler passes (never for a reported number)."""
    global TOTAL_LEN, N_OUT
    TOTAL_LEN, N_OUT = n, n - PROMPT_LEN


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def make_prompts(n, seed=1):
    rng = np.random.default_rng(seed)
    ids = rng.integers(3, 32000, (n, PROMPT_LEN))
    ids[:, 0] = 1  # BOS
    return ids.astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def all_blas_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        with all_blas_threads():
            return max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cheap_weights(args, hidden, n_layers):
    """Host weights in the reference layout for CPU TIMING only (values tiled from one small random block:
    dense BLAS time does not depend on them; generating 6 GB of normals for an 8B-shaped baseline would take
    longer than the baseline itself).  `n_layers` layers are materialised."""
    from dataclasses import replace
    from llama3_np_b200.synth import weight_shapes
    rng = np.random.default_rng(7)
    block = rng.standard_normal(1 << 20, dtype=np.float32)
    out = {}
    for key, shape, kind in weight_shapes(replace(args, n_layers=n_layers), hidden):
        n = int(np.prod(shape))
        w = np.resize(block, n).reshape(shape)
        if kind == "norm":
            w = 1.0 + 0.1 * w
        else:
            w = w * (0.5 if kind == "embed" else 0.85 / np.sqrt(shape[1]))
        out[key] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def oracle_decode_sample(n_prompts, positions):
    """stories15M through the oracle port: one decode step per entry of `positions` (each attends the whole
    cache up to its position, exactly as the reference's step at that position does - llama3.py:184-207).
    Returns (generated tokens per second, seconds)."""
    import llama3_np_b200  # noqa: F401
    from llama3_np_b200.config import named_config
    from llama3_np_b200.synth import make_weights
    from oracle import ref_llama3 as orc
    args, hidden = named_config("stories15M", max_batch_size=n_prompts)
    m = orc.OracleLlama(make_weights(args, hidden, seed=0), args)
    ids = make_prompts(n_prompts).astype(np.int64)[:, :1]
    with all_blas_threads():
        t0 = time.perf_counter()
        for p in positions:
            ids = m(ids, int(p))[:, -1, :].argmax(-1, keepdims=True)
        dt = time.perf_counter() - t0
    return n_prompts * len(positions) / dt, dt


def spread_positions(n):
    """n decode positions spread evenly over [PROMPT_LEN + 1, TOTAL_LEN - 1]: their mean is the mean position of the
    GPU arm's 247 decode steps, so both arms pay the same average attention cost."""
    return [int(round(x)) for x in np.linspace(PROMPT_LEN + 1, TOTAL_LEN - 1, n)]


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, ntok = a.prompts, 12
    for _ in range(a.warmup):
        oracle_decode_sample(B, spread_positions(2))
    times = []
    for _ in range(a.steps):
        _, dt = oracle_decode_sample(B, spread_positions(ntok))
        times.append(dt)
    value = B * ntok * a.steps / sum(times)
    cores = blas_threads()
    sample = (f"{B} prompts x {ntok} decode steps per bench step at positions spread over {PROMPT_LEN + 1}..{TOTAL_LEN - 1} "
              f"(mean {np.mean(spread_positions(ntok)):.0f}, as the GPU arm's 247 steps), NumPy float64-activation path, "
              f"BLAS threads {cores} of {os.cpu_count()} cpus")
    print(json.dumps({
        "impl": "reference", "metric": "decode tokens/s (batched greedy decode, generated tokens only)",
        "value": value, "unit": "tokens/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * sum(times) / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(a, B),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def config_dict(a, B):
    return {"workload": "stories15M batched greedy decode (BASELINE.json configs[1])",
            "prompts_per_gpu": B, "prompt_len": PROMPT_LEN, "total_len": TOTAL_LEN,
            "generated_per_prompt": N_OUT, "weights": "random-init, reference .npz layout, seed 0",
            "parallelism": f"dp{a.gpus} (independent prompts, no collective)",
            "l2": "working set (KV cache 906 MB fp32 at B=256) exceeds the 126 MB L2; L2 also flushed between steps"}


# ------------------------------------------------------------------------------------ shape records (extra / tp)
def dev_buf(m, nbytes):
    p = C.c_void_p()
    from llama3_np_b200 import _cabi
    _cabi.check(m._lib.l3_dev_alloc(m._h, nbytes, C.byref(p)), m._h)
    return p


def timed(m, fn, iters=1, reduce=None):
    from llama3_np_b200 import _cabi
    ms = C.c_float()
    _cabi.check(m._lib.l3_timer_start(m._h), m._h)
    for _ in range(iters):
        fn()
    _cabi.check(m._lib.l3_timer_stop(m._h, C.byref(ms)), m._h)
    v = ms.value / iters
    return reduce(v) if reduce else v


def shape_record(name, shape, dtype, B, L, n_decode, device=0, ids=None, iters=2, tp=None, e2e=True, workload=""):
    """One model shape on the device: prefill of [B, L] (+ argmax) and, if n_decode > 1, the greedy generate of
    n_decode tokens; device-resident ids for `value`, host ids through the Python surface for `e2e`."""
    import llama3_np_b200  # noqa: F401
    from llama3_np_b200 import Llama, _cabi
    from llama3_np_b200.config import named_config
    from llama3_np_b200.synth import make_weights, param_count
    hbm, tf, how = peaks()
    args, hidden = named_config(shape, max_batch_size=B, max_seq_len=L + max(n_decode, 1) + 2, dtype=dtype)
    kw = dict(device=device)
    red = None
    if tp:
        kw.update(tp_rank=tp["rank"], tp_world=tp["world"], tp_unique_id=tp["uid"])
        red = tp["max"]
    if shape == "stories15M":  # small enough for host weights in the reference layout
        m = Llama(make_weights(args, hidden, seed=0), args, **kw)
    else:
        m = Llama(None, args, hidden_dim=hidden, random_seed=0, **kw)
    lib, h = m._lib, m._h
    if ids is None:
        ids = np.random.default_rng(2).integers(3, args.vocab_size, (B, L))
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    d_ids = dev_buf(m, ids.nbytes)
    _cabi.check(lib.l3_memcpy_h2d(h, d_ids, ids.ctypes.data_as(C.c_void_p), ids.nbytes), h)
    d_out = dev_buf(m, B * max(n_decode, 1) * 8)
    wb = 4 if dtype == "float32" else 2
    nkv = args.n_heads if args.n_kv_heads is None else args.n_kv_heads
    hd = args.dim // args.n_heads
    D, FD, VS, NL = args.dim, hidden, args.vocab_size, args.n_layers
    G = tp["world"] if tp else 1
    clocks = ClockSampler(device).start()
    rec = {"name": name, "workload": workload, "shape": shape, "dtype": "f32" if dtype == "float32" else "bf16", "B": B,
           "prompt_len": L, "decode_tokens": n_decode, "n_layers": NL, "weights": "random-init (device RNG), seed 0", "tp": G}

    def prefill():
        _cabi.check(lib.l3_forward_dev(h, d_ids, B, L, 0, None, d_out), h)
    prefill()
    m.sync()
    ms_pf = timed(m, prefill, iters, red)
    flops = 2 * B * L * NL * (2 * D * D + 2 * D * nkv * hd + 3 * D * FD) + B * NL * (4 * L * L * D) // 2 + 2 * B * VS * D
    rec["prefill"] = {"ms": ms_pf, "tokens_per_s": B * L / (ms_pf / 1e3),
                      "roofline": {"bound": "tensor", "achieved": flops / (ms_pf / 1e3) / 1e12 / G, "peak": tf,
                                   "unit": "TFLOP/s per GPU", "frac": flops / (ms_pf / 1e3) / 1e12 / G / tf,
                                   "peak_source": how, "algorithmic_flops": flops}}
    tokens = None
    if n_decode > 1:
        total = L + n_decode

        def gen():
            _cabi.check(lib.l3_generate_greedy_dev(h, d_ids, B, L, total, d_out), h)
        gen()
        m.sync()
        m.launch_count(reset=True)
        ms_gen = timed(m, gen, iters, red)
        launches = m.launch_count(reset=True) // iters
        tokens = np.empty((B, n_decode), np.int64)
        _cabi.check(lib.l3_memcpy_d2h(h, tokens.ctypes.data_as(C.c_void_p), d_out, tokens.nbytes), h)
        ms_dec = (ms_gen - ms_pf) / (n_decode - 1)            # per decode step (the first token comes from the prefill)
        params = param_count(args, hidden) - VS * D + D        # weights read per decoded token
        pos_mid = L + n_decode // 2
        bytes_step = params * wb + B * NL * nkv * pos_mid * hd * 2 * wb
        rec.update(value=B / (ms_dec / 1e3), unit="tokens/s", ms_per_step=ms_dec, gpu_launches_per_generate=launches,
                   roofline={"bound": "hbm", "achieved": bytes_step / (ms_dec / 1e3) / 1e9 / G, "peak": hbm,
                             "unit": "GB/s per GPU", "frac": bytes_step / (ms_dec / 1e3) / 1e9 / G / hbm, "peak_source": how,
                             "algorithmic_bytes_per_step": bytes_step,
                             "note": "weights once per step + K/V of the mean position; whole decode step, all kernels"})
    else:
        rec.update(value=rec["prefill"]["tokens_per_s"], unit="tokens/s (prefill)", ms_per_step=ms_pf,
                   roofline=rec["prefill"]["roofline"])
    if e2e:
        # the call a user makes: host ids in; logits [B, 1, VS] float64 back (prefill) / every token read back (decode)
        t0 = time.perf_counter()
        logits = m(ids, 0)
        t_call = time.perf_counter() - t0
        rec["prefill"]["e2e"] = {"ms": t_call * 1e3, "tokens_per_s": B * L / t_call, "api": "Llama.__call__ (host ids, float64 logits back)",
                                 "h2d_bytes": int(ids.nbytes), "d2h_bytes": int(B * VS * 4)}
        del logits
        if n_decode > 1:
            t0 = time.perf_counter()
            out = [t for t in m.generate(ids, L + n_decode)]
            t_gen = time.perf_counter() - t0
            t_first = rec["prefill"]["ms"] / 1e3
            rec["e2e"] = {"value": B * (n_decode - 1) / max(t_gen - t_first, 1e-9), "unit": "tokens/s",
                          "ms_total": t_gen * 1e3, "api": "Llama.generate (lazy generator, one D2H per yielded step)",
                          "h2d_bytes_per_step": int(ids.nbytes), "d2h_bytes_per_step": int(B * n_decode * 8),
                          "tokens_equal_device_loop": bool(np.array_equal(np.concatenate(out, axis=1), tokens))}
        else:
            rec["e2e"] = {"value": rec["prefill"]["e2e"]["tokens_per_s"], "unit": "tokens/s (prefill)",
                          "api": rec["prefill"]["e2e"]["api"], "h2d_bytes_per_step": int(ids.nbytes), "d2h_bytes_per_step": int(B * VS * 4)}
    rec["clocks"] = clocks.stop()
    lib.l3_dev_free(h, d_ids)
    lib.l3_dev_free(h, d_out)
    m.close()
    return rec, tokens


def cpu_shape_baseline(shape, B, L, n_decode_sample):
    """The oracle port at the SAME widths with 1 and 2 layers (+ LM head), timed on the host cores: per-layer and
    head costs follow from the two runs; the full-depth figure is their linear extrapolation (the
    reference re-casts every fp32 weight to fp64 per call - a full-depth 8B-shaped run needs 30 GB and minutes)."""
    from dataclasses import replace
    import llama3_np_b200  # noqa: F401
    from llama3_np_b200.config import named_config
    from oracle import ref_llama3 as orc
    args, hidden = named_config(shape, max_batch_size=B, max_seq_len=L + n_decode_sample + 2)
    w = cheap_weights(args, hidden, 2)
    ids = np.random.default_rng(2).integers(3, args.vocab_size, (B, L)).astype(np.int64)
    t = {}
    with all_blas_threads():
        for nl in (1, 2):
            m = orc.OracleLlama(w, replace(args, n_layers=nl))
            t0 = time.perf_counter()
            nxt = m(ids, 0)[:, -1, :].argmax(-1, keepdims=True)
            t_pf = time.perf_counter() - t0
            t0 = time.perf_counter()
            for i in range(n_decode_sample):
                nxt = m(nxt, L + 1 + i)[:, -1, :].argmax(-1, keepdims=True)
            t[nl] = (t_pf, (time.perf_counter() - t0) / max(n_decode_sample, 1))
            del m
    NL = args.n_layers
    pf_layer, dec_layer = t[2][0] - t[1][0], t[2][1] - t[1][1]
    pf_full = t[1][0] + (NL - 1) * pf_layer
    dec_full = t[1][1] + (NL - 1) * dec_layer
    cores = blas_threads()
    out = {"kind": "port", "cores": cores, "unit": "tokens/s",
           "sample": f"oracle/ref_llama3.py at the {shape} widths with 1 and 2 layers + LM head, B={B}: prefill of {L} tokens and "
                     f"{n_decode_sample} decode steps each; full depth ({NL} layers) extrapolated linearly in the layer count",
           "measured_2_layers": {"prefill_s": t[2][0], "decode_s_per_step": t[2][1]},
           "prefill_tokens_per_s_extrapolated": B * L / pf_full}
    if n_decode_sample:
        out["value"] = B / dec_full
        out["decode_s_per_step_extrapolated"] = dec_full
    else:
        out["value"] = B * L / pf_full
    return out


def run_extras(a, local):
    from llama3_np_b200.tokenizer import Tokenizer  # noqa: F401  (host tokenizer stays on the host; ids are fed directly)
    recs = []
    dream = np.array([[1, 76, 505, 263, 12561]])  # "I have a dream" through the reference tokenizer

    def guarded(fn, name):
        try:
            recs.append(fn())
        except Exception as e:  # a failing sub-record must not take the headline down
            recs.append({"name": name, "error": f"{type(e).__name__}: {e}"})

    def s15m_b1():
        rec, _ = shape_record("s15m_b1", "stories15M", "float32", 1, 5, 251, device=local, ids=dream, iters=3,
                              workload="BASELINE.json configs[0]: stories15M, 'I have a dream', batch 1, fp32, total 256 tokens")
        rec50, _ = shape_record("s15m_b1_50", "stories15M", "float32", 1, 5, 45, device=local, ids=dream, iters=3, e2e=True)
        rec["total_50_tokens"] = {"tokens_per_s_generated_only": rec50["value"], "e2e_tokens_per_s": rec50["e2e"]["value"],
                                  "reference_style_tokens_per_s": 50 / (rec50["e2e"]["ms_total"] / 1e3),
                                  "note": "README.md:20 quotes 33 tokens/s counting the 5 prompt tokens and the prefill"}
        if not a.no_cpu_baseline:
            import llama3_np_b200  # noqa: F401
            from llama3_np_b200.config import named_config
            from llama3_np_b200.synth import make_weights
            from oracle import ref_llama3 as orc
            args, hidden = named_config("stories15M")
            m = orc.OracleLlama(make_weights(args, hidden, seed=0), args)
            with all_blas_threads():
                t0 = time.perf_counter()
                n = sum(1 for _ in m.generate(dream.astype(np.int64), 50))
                dt = time.perf_counter() - t0
            rec["cpu_baseline"] = {"value": n / dt, "unit": "tokens/s", "cores": blas_threads(), "kind": "port",
                                   "sample": f"the whole configs[0] run: generate(ids, 50) -> {n} tokens in {dt:.1f} s"}
        return rec

    def big(name, shape, B, L, nd, workload, cpu_decode_sample):
        def f():
            rec, _ = shape_record(name, shape, "bfloat16", B, L, nd, device=local, workload=workload)
            if not a.no_cpu_baseline:
                rec["cpu_baseline"] = cpu_shape_baseline(shape, B, min(L, 256), cpu_decode_sample)
                if L > 256:
                    rec["cpu_baseline"]["sample"] += f" (prompt shortened to 256 of {L} tokens: the reference materialises L x L fp64 scores)"
            return rec
        return f

    guarded(s15m_b1, "s15m_b1")
    guarded(big("1b", "llama3.2-1b", 1, 2048, 256, "BASELINE.json configs[2]: Llama-3.2-1B-shaped, bf16, prefill 2048 + decode 256", 2), "1b")
    guarded(big("8b_b1", "llama3-8b", 1, 128, 256, "BASELINE.json configs[3] on one GPU: Llama-3-8B-shaped, bf16, batch 1, 128-token prompt + 256 decode", 2), "8b_b1")
    guarded(big("8b_b32", "llama3-8b", 32, 128, 256, "BASELINE.json configs[3] on one GPU: Llama-3-8B-shaped, bf16, batch 32", 1), "8b_b32")
    guarded(big("8b_prefill", "llama3-8b", 1, 2048, 1, "north_star: Llama-3-8B-shaped bf16 prefill of 2048 tokens", 0), "8b_prefill")
    return recs


def run_tp(a, rank, world, local, dist):
    """configs[3] / [4]: the 8B-shaped model tensor-parallel over all ranks of this job."""
    from llama3_np_b200 import dp
    recs = []
    cases = [("8b_tp_b1", 1, 128, 256), ("8b_tp_b32", 32, 128, 256)]
    if world == 8:
        cases.append(("8b_tp_prefill_32k", 1, 32768, 1))
    for name, B, L, nd in cases:
        try:
            uid = dp.tp_unique_id(dist)
            tp = {"rank": rank, "world": world, "uid": uid, "max": lambda v: dp.max_over_ranks(v, dist)}
            rec, tokens = shape_record(name, "llama3-8b", "bfloat16", B, L, nd, device=local, tp=tp, e2e=False, iters=2 if L < 4096 else 1,
                                       workload=f"BASELINE.json configs[{4 if L > 4096 else 3}]: Llama-3-8B-shaped, bf16, tensor parallel over {world} GPUs")
            if tokens is not None:
                every = [None] * world
                dist.all_gather_object(every, tokens[:, :16].tolist())
                rec["ranks_agree_on_tokens"] = all(e == every[0] for e in every)
                if rank == 0:
                    # the same model (device RNG weights are a function of the global index) on ONE GPU: ms per step for
                    # the efficiency, and how far the greedy streams agree (bf16 partial sums are rounded per rank, so
                    # the streams may part at a near-tie; fp32 mode is checked against the oracle in tests/test_tp_gpu.py)
                    one, tok1 = shape_record(name + "_tp1", "llama3-8b", "bfloat16", B, L, nd, device=local, e2e=False, iters=2)
                    same = (tokens == tok1)
                    first_diff = [int(np.argmin(r)) if not r.all() else int(r.size) for r in same]
                    rec["vs_tp1"] = {"tp1_ms_per_step": one["ms_per_step"], "tp1_tokens_per_s": one["value"],
                                     "speedup": one["ms_per_step"] / rec["ms_per_step"],
                                     "efficiency": one["ms_per_step"] / rec["ms_per_step"] / world,
                                     "tokens_equal_tp1": bool(same.all()),
                                     "common_prefix_tokens_min": int(min(first_diff)), "common_prefix_tokens_mean": float(np.mean(first_diff)),
                                     "of": int(nd),
                                     "note": "bf16 partial sums are rounded per rank; random-init weights have top-1 margins of ~1e-3 of the "
                                             "logit scale, so the greedy streams part after a few tokens; parity of the TP path is "
                                             "tests/test_tp_gpu.py (fp32: token-identical to the oracle, bf16: inside the error bar)",
                                     "tp1_prefill_ms": one["prefill"]["ms"]}
            recs.append(rec)
        except Exception as e:
            recs.append({"name": name, "error": f"{type(e).__name__}: {e}"})
        dist.barrier()
    return recs


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # barriers and max-over-ranks only: the data path has no collective

    import llama3_np_b200  # noqa: F401
    from llama3_np_b200 import Llama, _cabi
    from llama3_np_b200.config import named_config
    from llama3_np_b200.synth import make_weights, param_count

    B = a.prompts
    dtype = {"f32": "float32", "bf16": "bfloat16"}[a.dtype]
    args, hidden = named_config("stories15M", max_batch_size=B, dtype=dtype)
    w = make_weights(args, hidden, seed=0)
    m = Llama(w, args, device=local)
    lib, h = m._lib, m._h
    from llama3_np_b200 import dp
    ids = dp.shard_prompts(make_prompts(B * world, seed=1), rank, world)  # weak scaling: B prompts per GPU

    # device-resident inputs / outputs for `value`
    d_ids, d_out = C.c_void_p(), C.c_void_p()
    _cabi.check(lib.l3_dev_alloc(h, ids.nbytes, C.byref(d_ids)), h)
    _cabi.check(lib.l3_dev_alloc(h, B * N_OUT * 8, C.byref(d_out)), h)
    _cabi.check(lib.l3_memcpy_h2d(h, d_ids, ids.ctypes.data_as(C.c_void_p), ids.nbytes), h)

    def dev_step():
        _cabi.check(lib.l3_flush_l2(h), h)
        _cabi.check(lib.l3_generate_greedy_dev(h, d_ids, B, PROMPT_LEN, TOTAL_LEN, d_out), h)

    def barrier():
        m.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(ms):
        return dp.max_over_ranks(ms, dist)

    for _ in range(a.warmup):
        dev_step()
    barrier()
    m.launch_count(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = C.c_float()
    _cabi.check(lib.l3_timer_start(h), h)
    for _ in range(a.steps):
        dev_step()
    _cabi.check(lib.l3_timer_stop(h, C.byref(ms)), h)
    launches = m.launch_count(reset=True)
    barrier()
    dev_ms = max_over_ranks(ms.value)
    tokens_dev = np.empty((B, N_OUT), np.int64)
    _cabi.check(lib.l3_memcpy_d2h(h, tokens_dev.ctypes.data_as(C.c_void_p), d_out, tokens_dev.nbytes), h)

    # e2e: the drop-in generator, host ids in, every token read back
    def e2e_step():
        _cabi.check(lib.l3_flush_l2(h), h)
        out = [t for t in m.generate(ids, TOTAL_LEN)]
        return np.concatenate(out, axis=1)

    for _ in range(max(1, a.warmup // 2)):
        tokens_e2e = e2e_step()
    barrier()
    _cabi.check(lib.l3_timer_start(h), h)
    for _ in range(a.steps):
        tokens_e2e = e2e_step()
    _cabi.check(lib.l3_timer_stop(h, C.byref(ms)), h)
    e2e_ms = max_over_ranks(ms.value)
    clk = clocks.stop() if rank == 0 else None
    assert np.array_equal(tokens_e2e, tokens_dev), "device loop and generator disagree"

    tokens_per_step_job = B * N_OUT * world
    value = tokens_per_step_job * a.steps / (dev_ms / 1e3)
    e2e_value = tokens_per_step_job * a.steps / (e2e_ms / 1e3)

    line = None
    if rank == 0:
        hbm, tf, how = peaks()
        wb = 4 if a.dtype == "f32" else 2
        nl, hn, hd, D, FD, VS = args.n_layers, args.n_heads, args.dim // args.n_heads, args.dim, hidden, args.vocab_size
        pos_mid = (PROMPT_LEN + TOTAL_LEN) // 2
        step_ms = dev_ms / a.steps / N_OUT  # one decode step of the whole batch (the prefill is 1 of 248 steps)
        # Dominant kernel of a decode step: decode_stack_kernel (every layer of the step, one launch), ~88 % of the
        # step's device time (profiles/r02_launches_decode_step.csv); the LM head + finalize are the rest.
        # Its algorithmic bytes per launch: the layers' weights once (each of the ceil(B / 12) clusters
        # streams them, but from L2) + K and V of every cached position of every sequence and layer.
        stack_ms, lm_ms = C.c_float(), C.c_float()
        _cabi.check(lib.l3_bench_kernel(h, 4, B, pos_mid, 30, C.byref(stack_ms)), h)
        _cabi.check(lib.l3_bench_kernel(h, 5, B, pos_mid, 30, C.byref(lm_ms)), h)
        layer_params = nl * (4 * D * D + 3 * D * FD + 2 * D) + D
        kv_bytes = B * nl * hn * pos_mid * hd * 2 * wb
        stack_bytes = layer_params * wb + kv_bytes + B * D * wb * 3      # + embedding rows in, normalised rows (hi, lo) out
        traffic, traffic_note = None, None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of the same kernel and shape
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                t = json.load(f)["decode_stack_kernel"]
            if B == 256 and a.dtype == "f32" and TOTAL_LEN == 256:
                traffic, traffic_note = t["traffic_bytes_per_launch"], t["config"] + "; " + t["source"]
        except Exception:
            pass
        achieved = stack_bytes / (stack_ms.value / 1e3) / 1e9
        params = param_count(args, hidden) - VS * D + D
        step_bytes = sum(params * wb + B * nl * hn * (p + 1) * hd * 2 * wb for p in range(PROMPT_LEN + 1, TOTAL_LEN))
        roofline = {"bound": "hbm", "kernel": "decode_stack_kernel (all 6 layers of one decode step for 256 sequences, one launch)",
                    "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                    "traffic_note": traffic_note, "peak_source": how, "algorithmic_work_per_launch": stack_bytes,
                    "launch_ms": stack_ms.value, "at": f"B={B}, position {pos_mid}",
                    "note": "timed alone with CUDA events on the library stream, L2 flushed between launches; the kernel is bound by "
                            "FFMA issue + cluster exchanges, not by HBM (DESIGN.md 4.0): the fraction says how far the step is from "
                            "the data-movement floor",
                    "per_decode_step_ms": {"decode_stack_kernel": stack_ms.value, "lm_head (gemm_tc_kernel, 256-wide tiles, fused argmax)": lm_ms.value},
                    "decode_step_ms_measured": step_ms,
                    "whole_step_hbm_frac": step_bytes / (dev_ms / a.steps / 1e3) / 1e9 / hbm}
        line = {
            "metric": "decode tokens/s (batched greedy decode, generated tokens only)",
            "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic", "config": config_dict(a, B),
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(ids.nbytes),
                    "d2h_bytes_per_step": int(B * N_OUT * 4), "ms_per_step": e2e_ms / a.steps,
                    "api": "Llama.generate (lazy generator, one D2H per yielded step)"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "env_switches": {k: os.environ[k] for k in ENV_SWITCHES if k in os.environ},
        }
    _cabi.check(lib.l3_dev_free(h, d_ids), h)
    _cabi.check(lib.l3_dev_free(h, d_out), h)
    m.close()

    if rank == 0:
        if a.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": blas_threads(), "kind": "port", "sample": "skipped (--no-cpu-baseline)"}
        else:
            n_cpu, ntok = min(B, 256), 24
            cpu_val, cpu_dt = oracle_decode_sample(n_cpu, spread_positions(ntok))
            cores = blas_threads()
            line["cpu_baseline"] = {"value": cpu_val, "unit": "tokens/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_cpu} prompts x {ntok} decode steps at positions spread over {PROMPT_LEN + 1}..{TOTAL_LEN - 1} "
                                              f"({cpu_dt:.1f} s of oracle/ref_llama3.py, BLAS threads {cores})"}
    if world == 1 and not a.no_extra:
        line["extra"] = run_extras(a, local)
    if world > 1 and not a.no_extra:
        tp = run_tp(a, rank, world, local, dist)
        if rank == 0:
            line["tp"] = tp
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--prompts", type=int, default=256, help="prompts per GPU")
    ap.add_argument("--total-len", type=int, default=256, help="profiling only: shorter generate")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling only: skip the CPU legs")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip the `extra` / `tp` sub-records")
    a = ap.parse_args()
    if a.total_len != 256:
        set_total_len(a.total_len)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
