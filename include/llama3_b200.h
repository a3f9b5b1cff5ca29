/* llama3_b200.h - C-ABI of the B200-native (sm_100a) Llama-3 forward / greedy-generate path.
 *
 * The reference (swap357/llama3.np) is pure Python + NumPy and has no FFI of its own; the
 * boundary it exposes is its Python surface.  Every entry point below names the reference
 * behaviour it replaces (file:line relative to the reference checkout).  The Python host
 * (`llama3.np_b200/llama3.py`) binds these with ctypes and mirrors `Llama` / `ModelArgs`.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative L3_E* code on failure, with a message retrievable through l3_last_error().
 * One handle = one CUDA device + one stream; a handle is not re-entrant.  All host
 * buffers may be pageable; "dev" variants take device pointers and never touch the host.
 */
#ifndef LLAMA3_B200_H
#define LLAMA3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L3_OK 0
#define L3_EINVAL (-1)   /* bad argument (maps to ValueError)              */
#define L3_ECUDA (-2)    /* CUDA runtime / launch failure (RuntimeError)    */
#define L3_ESTATE (-3)   /* call out of order (e.g. forward before finalize) */
#define L3_ENOMEM (-4)
#define L3_ENCCL (-5)

#define L3_DTYPE_F32 0   /* fp32 weights, activations, KV cache: token-identical mode */
#define L3_DTYPE_BF16 1  /* bf16 weights / KV cache / GEMM operands, fp32 accumulation  */

/* ModelArgs (reference config.py:5-19) + the FFN width the reference infers from the
 * up_proj weight shape (llama3.py:86-95).  rope base is fixed at 10000 (llama3.py:31). */
typedef struct L3Config {
  int32_t dim;
  int32_t n_layers;
  int32_t n_heads;
  int32_t n_kv_heads;     /* already resolved: n_heads when ModelArgs.n_kv_heads is None */
  int32_t vocab_size;
  int32_t max_seq_len;
  int32_t max_batch_size;
  int32_t hidden_dim;
  float norm_eps;
  int32_t dtype;          /* L3_DTYPE_* */
  int32_t device;         /* CUDA ordinal */
  int32_t tp_rank;        /* tensor-parallel rank / world (1 = single GPU) */
  int32_t tp_world;
  int32_t flags;          /* L3_FLAG_* */
} L3Config;

#define L3_FLAG_NO_GRAPH 1      /* never capture CUDA graphs (debug)                      */
#define L3_FLAG_NO_TENSORCORE 2 /* keep SIMT GEMMs (A/B against the tcgen05 kernels)         */
#define L3_FLAG_NO_PDL 4        /* process-wide: no programmatic dependent launch           */
#define L3_FLAG_NO_MEGA 8       /* batch-1 decode as one kernel per projection (A/B against
                                   the persistent single-kernel step)                       */

typedef struct L3Model L3Model;

/* Library-level (no handle): version string, and the last error of a failed l3_create. */
const char* l3_version(void);
const char* l3_last_error(const L3Model* m); /* m may be NULL */
int l3_device_count(int* out);

/* -- lifetime: replaces Llama.__init__ (llama3.py:265-283) -------------------------- */
int l3_create(const L3Config* cfg, L3Model** out);
/* Upload one tensor of the reference's .npz key layout (llama3.py:219-235, 269, 280-281):
 * float32, C-contiguous, stored [out, in].  Keys are the reference's HF-style names. */
int l3_load_weight(L3Model* m, const char* key, const float* host, const int64_t* shape, int ndim);
/* Random-init on the device for shapes that are impractical to build on the host
 * (8B-shaped benches).  Same scales as synth.make_weights, different stream of numbers. */
int l3_fill_random(L3Model* m, uint64_t seed);
/* Packed device-layout weight cache (replaces a repeat of np.load + per-tensor packing,
 * llama3.py:269 / utils.py:4-5, on the second and later starts).  l3_save_packed writes the weight matrices
 * exactly as they sit on the device after the l3_load_weight calls (fused q|k|v rows, interleaved gate/up rows,
 * this rank's tensor-parallel slices, model dtype) to `path` (atomically: temp file + rename) together with
 * `source_digest`, the caller's hex digest of the checkpoint they came from.  l3_load_packed, called INSTEAD of
 * the l3_load_weight calls and before l3_finalize, streams such a file into the device buffers; it fails with
 * L3_EINVAL - leaving the model unloaded - if shape, dtype, tensor-parallel placement, digest or any tensor
 * checksum differ.  l3_packed_info reads a file's header without a model (hidden_dim is needed for l3_create). */
int l3_save_packed(L3Model* m, const char* path, const char* source_digest);
int l3_load_packed(L3Model* m, const char* path, const char* source_digest /* NULL: accept any */);
int l3_packed_info(const char* path, L3Config* cfg_out, char* digest_out, int digest_cap);
/* RoPE tables as computed by compute_cos_sin_cache in float64 on the host
 * (llama3.py:31-38), [max_seq_len, head_dim/2] each. */
int l3_set_rope_tables(L3Model* m, const double* cos_tab, const double* sin_tab);
/* Pack (fused QKV rows, interleaved gate/up rows, optional bf16 copies), allocate and
 * zero the KV caches (llama3.py:138-153) and the activation workspace. */
int l3_finalize(L3Model* m);
int l3_destroy(L3Model* m);
/* Zero the KV caches (the reference never does; a new Llama instance starts at zero). */
int l3_reset_cache(L3Model* m);
/* Tensor parallel: hand the handle an initialised NCCL communicator (ncclComm_t) made by
 * the caller, or let it create one from a 128-byte ncclUniqueId. */
int l3_tp_init(L3Model* m, const void* nccl_unique_id_128);
int l3_nccl_unique_id(void* out_128);

/* -- the forward step: replaces Llama.__call__ (llama3.py:285-308) ------------------- */
/* ids [B, L] int32 row-major (host).  logits_out: [B, vocab] float32 or NULL.
 * argmax_out: [B] int64 (first maximum wins, llama3.py:320) or NULL. */
int l3_forward(L3Model* m, const int32_t* ids, int B, int L, int start_pos,
               float* logits_out, int64_t* argmax_out);
/* Same, all pointers on the device; enqueues on the handle's stream and returns. */
int l3_forward_dev(L3Model* m, const int32_t* d_ids, int B, int L, int start_pos,
                   float* d_logits_out, int64_t* d_argmax_out);

/* -- the greedy loop: replaces Llama.generate (llama3.py:310-321) -------------------- */
/* Runs prefill at pos 0 and then decode steps with pos = L + i (the reference's schedule,
 * including its skipped slot L).  out: [B, max_new_tokens - L] int64 row-major.
 * Returns L3_EINVAL when max_new_tokens > max_seq_len or L >= max_new_tokens yields nothing
 * (n_out = 0 is allowed and writes nothing). */
int l3_generate_greedy(L3Model* m, const int32_t* ids, int B, int L, int max_new_tokens,
                       int64_t* out);
int l3_generate_greedy_dev(L3Model* m, const int32_t* d_ids, int B, int L, int max_new_tokens,
                           int64_t* d_out);
/* Incremental form used by the lazy Python generator: start enqueues the prefill,
 * each next() runs one step and copies that step's [B] token ids to the host. */
int l3_generate_begin(L3Model* m, const int32_t* ids, int B, int L);
/* pos_offset 0: decode step i at pos = L + i (Llama.generate, llama3.py:316-318);
 * pos_offset -1: pos = L + i - 1 (llama_generate of llama3_simple.py:272-280). */
int l3_generate_begin_ex(L3Model* m, const int32_t* ids, int B, int L, int pos_offset);
int l3_generate_next(L3Model* m, int64_t* out_B);
/* Extension (the reference generates equal-length prompts only and checks EOS for row 0 in the caller,
 * llama3.py:341-343): prompts of different lengths in one batch.  ids [B, Lmax] int32, right-padded
 * (padding values ignored); lens [B]; out [B, max_new_tokens] int64.  Sequence b yields exactly what it
 * would yield alone: Llama.generate(ids_b, lens[b] + max_new_tokens) for pos_offset 0, llama_generate for
 * pos_offset -1.  eos_id >= 0: a sequence that emitted eos_id keeps emitting it. */
int l3_generate_ragged(L3Model* m, const int32_t* ids, const int32_t* lens, int B, int Lmax, int max_new_tokens,
                       int pos_offset, int eos_id, int64_t* out);

/* -- state inspection (tests): the layer's caches in the reference's layout
 * [max_batch, max_seq_len, n_kv_heads, head_dim] (llama3.py:138-153), as float32. */
int l3_read_cache(L3Model* m, int layer, float* k_out, float* v_out);

/* -- per-op entry points (parity tests against the reference's functions) ------------
 * All take HOST float32 buffers, run the same kernels the model path launches. */
/* RMSNorm.__call__ (llama3.py:111-114): x [rows, dim] */
int l3_op_rmsnorm(int device, const float* x, const float* w, float eps, int rows, int dim, float* out);
/* y = x @ W.T with W [n, k] (the `x @ self.*_weight` lines, llama3.py:99-102,166-168,211,307).
 * path: 0 auto, 1 row-streaming GEMV, 2 SIMT tiled GEMM, 3 tcgen05 (bf16 or 3xTF32 operands),
 * 4 tcgen05 with swapped operand roles (rows <= 128, n >= 128).
 * w_bf16 != 0 rounds W (and for path 3 x) to bf16 first. */
int l3_op_linear(int device, const float* x, const float* w, int rows, int n, int k,
                 int path, int w_bf16, float* out);
/* apply_rotary_emb (llama3.py:41-76) on one tensor x [B, L, heads, head_dim] at positions
 * start_pos..start_pos+L-1; tables [max_pos, head_dim/2] float64. */
int l3_op_rope(int device, const float* x, const double* cos_tab, const double* sin_tab,
               int B, int L, int heads, int head_dim, int start_pos, float* out);
/* silu(gate) * up (llama3.py:27-28, 99-101) elementwise over n values */
int l3_op_swiglu(int device, const float* gate, const float* up, int64_t n, float* out);
/* Attention core (llama3.py:190-207): q [B, L, HN, HD]; k, v caches in the reference layout
 * [B, T, KVHN, HD] holding T = start_pos + L valid positions; out [B, L, HN*HD]. */
int l3_op_attention(int device, const float* q, const float* k, const float* v,
                    int B, int L, int n_heads, int n_kv_heads, int head_dim, int start_pos,
                    int kv_bf16 /* 0 fp32 cache, 1 bf16 cache, 2 bf16 tensor-core prefill (L > 1, head_dim 64/128),
                                   3 bf16 cache, decode through the exact (non tensor-core) kernels */,
                    int nsplit /* decode split-KV factor, 0 = auto */, float* out);
/* logits[:, -1, :].argmax(-1) (llama3.py:320): first maximum wins. */
int l3_op_argmax(int device, const float* logits, int rows, int n, int64_t* out);

/* Debug: switch the tcgen05 GEMM's clock64 milestone stamps on/off and read the last 64
 * stamps of CTA (0,0) (indices documented in gemm_tc.cu). */
int l3_debug_tc_timeline(int device, int enable, uint64_t* out64);

/* -- measurement helpers (bench.py): CUDA events on the handle's own stream ---------- */
int l3_sync(L3Model* m);
int l3_timer_start(L3Model* m);
int l3_timer_stop(L3Model* m, float* ms_out);      /* records, synchronises, returns ms */
int l3_dev_alloc(L3Model* m, int64_t bytes, void** out);
int l3_dev_free(L3Model* m, void* p);
int l3_memcpy_h2d(L3Model* m, void* dst, const void* src, int64_t bytes);
int l3_memcpy_d2h(L3Model* m, void* dst, const void* src, int64_t bytes);
int l3_flush_l2(L3Model* m);                        /* writes a buffer larger than L2 */
/* Number of kernel launches (graph nodes count individually) since the last call. */
int l3_launch_count(L3Model* m, int64_t* out, int reset);
/* Time one kernel family in isolation on the current model state (bench roofline leg):
 * which: 0 = decode attention at position `pos` for batch B, 1 = LM head for B rows, 2 = the FFN leg of a
 * layer (RMSNorm + gate|up + down), 3 = the two residual projections of a layer (Wo, Wdown: one kernel symbol);
 * batched decode path (decode_stack.cu), each launch timed alone after an L2 flush: 4 = the cluster-resident kernel
 * (every layer of one decode step at `pos`), 5 = its LM head with the fused argmax.
 * Runs `iters` launches bracketed by events; returns average ms. */
int l3_bench_kernel(L3Model* m, int which, int B, int pos, int iters, float* avg_ms);

/* Debug: %globaltimer stamps of the last persistent decode step, [n_sm][512] (layout in decode_mega.cu);
 * needs L3_MEGA_DBG=1 in the environment when the model is created. */
int l3_debug_mega_timeline(L3Model* m, uint64_t* out, int64_t capacity);

/* Debug of the cluster-resident batched-decode kernel (decode_stack.cu), last step; needs L3_STACK_DBG=1 when the
 * model is created.  which 0: %globaltimer stamps uint64 [clusters * 8][128] (row = CTA, layout in the kernel);
 * which 1: float [n_layers][max_batch_size][dim], the residual stream after every layer. */
int l3_debug_stack(L3Model* m, int which, void* out, int64_t capacity_bytes);

/* Micro-benchmark of the row-streaming GEMV (y = W x, W [n, k]) at one shape, weights rotated
 * over more copies than fit the L2: average ms per launch over `iters` launches. */
int l3_bench_gemv(int device, int n, int k, int w_bf16, int rows, int iters, float* avg_ms);

/* Micro-benchmark of the tensor pipe alone (csrc/mma_probe.cu, scripts/mma_cost.py): cycles per
 * tcgen05.mma of M = 128, K = 32 bytes, width n (16..256), kind 0 = bf16 / 1 = tf32, issued by one thread
 * rotating over nacc TMEM accumulators (nacc * n <= 512), `iters` groups of four, on `ctas` CTAs at once.
 * cycles_per_mma[0] = issue loop, [1] = issue + completion (maximum over the CTAs).  No reference
 * counterpart: a design input for the GEMM kernels. */
int l3_probe_mma(int device, int kind, int n, int nacc, int iters, int ctas, double* cycles_per_mma);

#ifdef __cplusplus
}
#endif
#endif /* LLAMA3_B200_H */
