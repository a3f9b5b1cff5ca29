#!/usr/bin/env python
"""bench.py - BASELINE.json's headline metric on its configs[1]:

  stories15M (random-init weights in the reference's .npz layout), batched greedy decode of
  256 independent prompts per GPU (BOS + 7 ids each, `generate(ids, 256)` -> 248 tokens per
  prompt), prompts sharded data-parallel over the GPUs with no data-path collective.

A "step" is one full `generate` over the rank's batch (one prefill + 247 decode steps).
`value` = generated tokens/s of the whole job with the prompt ids already resident in HBM
(l3_generate_greedy_dev); `e2e` = the same through the drop-in `Llama.generate` generator
with host ids (H2D inside) and every yielded token read back to the host.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f32|bf16] [--impl reference]

Under torchrun (N > 1) each rank drives one GPU; timing is CUDA events on the library's own
stream, max over ranks.  `--impl reference` times the CPU oracle port of the reference
(oracle/ref_llama3.py; the reference is NumPy-only Python, nothing to compile) on a bounded
sample of the same workload.
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
def oracle_sample(n_prompts, n_tokens, threads_note=True):
    """Time the oracle port on `n_prompts` prompts for `n_tokens` generated tokens each."""
    import llama3_np_b200  # noqa: F401
    from llama3_np_b200.config import named_config
    from llama3_np_b200.synth import make_weights
    from oracle import ref_llama3 as orc
    args, hidden = named_config("stories15M", max_batch_size=n_prompts)
    w = make_weights(args, hidden, seed=0)
    m = orc.OracleLlama(w, args)
    ids = make_prompts(n_prompts).astype(np.int64)
    with all_blas_threads():
        t0 = time.perf_counter()
        n = 0
        for _ in m.generate(ids, PROMPT_LEN + n_tokens):
            n += 1
        dt = time.perf_counter() - t0
    return n_prompts * n / dt, dt


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


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, ntok = a.prompts, 12
    for _ in range(a.warmup):
        oracle_sample(B, 2)
    vals, times = [], []
    for _ in range(a.steps):
        v, dt = oracle_sample(B, ntok)
        vals.append(v)
        times.append(dt)
    total_tokens = B * ntok * a.steps
    value = total_tokens / sum(times)
    cores = blas_threads()
    sample = (f"{B} prompts x {ntok} of {N_OUT} generated tokens per step (positions {PROMPT_LEN}..{PROMPT_LEN + ntok}), "
              f"NumPy float64-activation path, BLAS threads {cores} of {os.cpu_count()} cpus")
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


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
        return dp.max_over_ranks(ms, dist, device="cuda" if dist is not None else None)

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

    if rank == 0:
        hbm, tf, how = peaks()
        wb = 4 if a.dtype == "f32" else 2
        # --- roofline of the dominant kernel family, timed in isolation at the mean position
        pos_mid = (PROMPT_LEN + TOTAL_LEN) // 2
        t_attn, t_lm, t_ffn = C.c_float(), C.c_float(), C.c_float()
        _cabi.check(lib.l3_bench_kernel(h, 0, B, pos_mid, 60, C.byref(t_attn)), h)
        _cabi.check(lib.l3_bench_kernel(h, 1, B, pos_mid, 20, C.byref(t_lm)), h)
        _cabi.check(lib.l3_bench_kernel(h, 2, B, pos_mid, 60, C.byref(t_ffn)), h)
        nl, hn, hd, D, FD, VS = args.n_layers, args.n_heads, args.dim // args.n_heads, args.dim, hidden, args.vocab_size
        kv_bytes = B * hn * (pos_mid + 1) * hd * 2 * wb + 2 * B * D * 4          # K+V read, q read, ctx write
        lm_bytes = VS * D * wb + B * VS * 4 + B * D * 4                            # weights + logits write + x
        ffn_bytes = 3 * D * FD * wb + 2 * B * D * 4 + 2 * B * FD * 4
        # Per decode step and KERNEL SYMBOL: decode attention (one launch per layer), the residual-epilogue
        # GEMM (Wo and Wdown: two launches per layer of one symbol), the LM head; the FFN leg (RMSNorm + two
        # GEMMs, three symbols) is listed for the step breakdown only.
        t_res = C.c_float()
        _cabi.check(lib.l3_bench_kernel(h, 3, B, pos_mid, 60, C.byref(t_res)), h)
        res_flops = 2.0 * B * (D * D + D * FD)                                   # Wo + Wdown, algorithmic (not x3 for 3xTF32)
        lm_flops = 2.0 * B * VS * D
        sym = {
            "attn_decode_kernel": dict(step_ms=t_attn.value * nl, launch_ms=t_attn.value, bound="hbm", work=kv_bytes),
            "gemm_tc_kernel<residual epilogue> (Wo + Wdown)": dict(step_ms=t_res.value * nl, launch_ms=t_res.value / 2,
                                                                   bound="tensor", work=res_flops / 2),
            "gemm_tc_kernel<argmax epilogue> (LM head)": dict(step_ms=t_lm.value, launch_ms=t_lm.value, bound="tensor", work=lm_flops),
        }
        dom = max(sym, key=lambda k: sym[k]["step_ms"])   # the kernel symbol with the most time per decode step
        d = sym[dom]
        traffic, traffic_note = None, None
        try:  # DRAM bytes per launch from the committed ncu --set full capture of the same kernel and shape
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                t = json.load(f).get(dom.split(" ")[0])
            if t and B == 256 and a.dtype == "f32" and TOTAL_LEN == 256:
                traffic = t["traffic_bytes_per_launch"]
                traffic_note = t["config"] + "; " + t["source"]
        except Exception:
            pass
        if d["bound"] == "hbm":
            achieved, peak, unit = d["work"] / (d["launch_ms"] / 1e3) / 1e9, hbm, "GB/s"
        else:
            achieved, peak, unit = d["work"] / (d["launch_ms"] / 1e3) / 1e12, tf, "TFLOP/s"
        roofline = {"bound": d["bound"], "kernel": dom, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": how,
                    "algorithmic_work_per_launch": d["work"], "launch_ms": d["launch_ms"],
                    "at": f"B={B}, position {pos_mid}",
                    "note": ("M = 256 projections of a 288-wide model: 18-96 tiles of 9-24 k-blocks each, bound by the "
                             "launch -> TMA -> MMA -> epilogue latency chain, not by HBM or tensor throughput (DESIGN.md 6); "
                             "fp32 mode runs them as 3xTF32 (three TF32 MMAs per product)") if d["bound"] == "tensor" else None,
                    "per_symbol": {k: {"step_ms": v["step_ms"], "launch_ms": v["launch_ms"], "bound": v["bound"],
                                       "frac": (v["work"] / (v["launch_ms"] / 1e3) / (1e9 * hbm if v["bound"] == "hbm" else 1e12 * tf))}
                                   for k, v in sym.items()},
                    "per_decode_step_ms": {"attn_decode_kernel": t_attn.value * nl, "lm_head (gemm_tc_kernel)": t_lm.value,
                                           "ffn (rmsnorm + 2 gemm_tc_kernel)": t_ffn.value * nl,
                                           "wo + w2 (gemm_tc_kernel, residual epilogue)": t_res.value * nl},
                    "decode_step_ms_measured": dev_ms / a.steps / (N_OUT)}
        # whole-step algorithmic HBM bytes: weights once per decode step + KV read per position
        params = param_count(args, hidden) - VS * D + D
        step_bytes = sum(params * wb + B * nl * hn * (p + 1) * hd * 2 * wb for p in range(PROMPT_LEN + 1, TOTAL_LEN))
        roofline["whole_step_hbm_frac"] = step_bytes / (dev_ms / a.steps / 1e3) / 1e9 / hbm

        n_cpu, tok_cpu = min(B, 256), 48
        if a.no_cpu_baseline:
            cpu_val, cpu_dt = None, 0.0
        else:
            cpu_val, cpu_dt = oracle_sample(n_cpu, tok_cpu)
        cores = blas_threads()
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
            "cpu_baseline": {"value": cpu_val, "unit": "tokens/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} prompts x {tok_cpu} of {N_OUT} generated tokens "
                                       f"({cpu_dt:.1f} s of oracle/ref_llama3.py, BLAS threads {cores})"},
        }
        print(json.dumps(line))
    _cabi.check(lib.l3_dev_free(h, d_ids), h)
    _cabi.check(lib.l3_dev_free(h, d_out), h)
    m.close()
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
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling only: skip the CPU leg")
    a = ap.parse_args()
    if a.total_len != 256:
        set_total_len(a.total_len)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
