// Shared device helpers and the internal kernel-launch interface of libllama3_b200.so.
// Everything here is sm_100a-only product code; the public boundary is include/llama3_b200.h.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define L3_WARP 32
#define L3_FULL 0xffffffffu

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of a step is launched with cudaLaunchAttributeProgrammaticStreamSerialization:
// it may become resident while its predecessor is still running, does its private prologue
// (barrier init, TMEM allocation, descriptor prefetch, smem carve-up) and then blocks in
// pdl_wait() until the predecessor grid has completed and its writes are visible.
// pdl_launch() at kernel entry lets the successor be scheduled as early as possible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern bool g_l3_pdl;  // l3_api.cu; false = plain stream-ordered launches
extern bool g_l3_pdl_next;  // l3_api.cu; set right before ONE launch that should overlap its predecessor's tail (consumed by launch_k)
// Pins a kernel to the max-shared-memory carveout (once per function and device).  The tcgen05
// GEMM needs ~200 KB of shared memory; if its neighbours in the stream ran with the default
// carveout every GEMM launch would pay an SM drain + L1/shared reconfiguration.
bool l3_carveout_seen(const void* fn);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                   Args... args) {
  if (!l3_carveout_seen((const void*)kern))
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (g_l3_pdl || g_l3_pdl_next) ? 1 : 0;
  g_l3_pdl_next = false;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ------------------------------------------------------------------ numeric helpers
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// Exact split of an fp32 value for the 3xTF32 tensor-core scheme (gemm_tc.cu): hi keeps the
// sign, exponent and top 10 mantissa bits (a valid TF32 number), lo = v - hi is exact in fp32.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  lo = v - hi;
}

// silu exactly as the reference writes it: x * (1 / (1 + exp(-x)))   (llama3.py:27-28)
__device__ __forceinline__ float silu_ref(float x) { return x * (1.0f / (1.0f + expf(-x))); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(L3_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(L3_FULL, v, o));
  return v;
}

// 16-byte streaming load that does not pollute L1 (weights / KV are read once per step).
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// Whole-range L2 prefetch (bytes: multiple of 16, 16-byte aligned address): one instruction, no
// register or shared-memory destination.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Unpack helpers: VEC consecutive elements of type T starting at a 16B/8B aligned address.
// float: 4 per 16 B.  bf16: 8 per 16 B.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y);
    v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    uint4 r = ldg_stream16(p);
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y);
    v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 r = ldg_stream16(p);
    uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// ------------------------------------------------------------------ epilogue description
// What a projection kernel does with each pair of adjacent outputs (col, col+1) of row m.
enum Epi : int {
  EPI_STORE = 0,   // out[m, col] = v
  EPI_RESID = 1,   // out[m, col] = resid[m, col] + v          (llama3.py:253, 259)
  EPI_SWIGLU = 2,  // out[m, col/2] = silu(v_even) * v_odd      (llama3.py:99-101; W rows interleaved gate/up)
  EPI_ROPE_KV = 3, // rotate q/k pairs (llama3.py:41-76), append k, v to the cache (llama3.py:184-185)
  EPI_ARGMAX = 4   // LM head without materialised logits: best[m] = max over columns (llama3.py:320)
};

// Order-preserving 64-bit key for the fused argmax: larger value wins, then the SMALLER index
// (NumPy's first-maximum rule), so a plain unsigned atomicMax merges partial results.
__device__ __forceinline__ unsigned long long argmax_key(float v, int idx) {
  uint32_t b = __float_as_uint(v);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

struct EpiArgs {
  float* out;            // fp32 destination (STORE / RESID / SWIGLU / q of ROPE_KV); may be null
  bf16* out_bf16;        // optional bf16 mirror of `out` for tensor-core consumers; may be null
  float* out_lo;         // if set (STORE / SWIGLU): `out` receives the TF32 hi part, out_lo the lo part
  unsigned long long* best;  // ARGMAX: [rows] packed (value, index) keys, zero = empty
  int col_offset;        // ARGMAX: global index of column 0 (vocab-sharded LM head)
  int ld_out;            // leading dimension of out / out_bf16
  const float* resid;    // RESID: source of the residual (may alias out)
  // ROPE_KV: fused row layout [q: HN*HD | k: KVHN*HD | v: KVHN*HD]
  void* cache_k;         // [maxB, KVHN, M, HD] of KVT (this layer)
  void* cache_v;
  const float* cos_tab;  // [M, HD/2]
  const float* sin_tab;
  const int* pos_ptr;    // device scalar: start_pos of this call
  const int* row_pos;    // ragged batches: per-sequence start position [B] (null: *pos_ptr for all)
  const int* row_len;    // ragged prefill: per-sequence prompt length [B]; tokens t >= row_len[b] are padding
                         // and leave the cache untouched (null: every token is real)
  int L;                 // tokens per sequence in this call: row m -> (b = m / L, t = m % L)
  int HD, HN, KVHN, M;   // head_dim, local heads, local kv heads, max_seq_len
};

template <typename KVT>
__device__ __forceinline__ void epilogue_pair(int epi, const EpiArgs& e, int m, int col, float v0, float v1,
                                              bool has1) {
  if (epi == EPI_STORE) {
    const size_t o = (size_t)m * e.ld_out + col;
    const bool vec = has1 && !(e.ld_out & 1);  // col is even: 8-byte aligned pair
    if (e.out_lo) {
      float h0, l0, h1, l1;
      split_tf32(v0, h0, l0);
      split_tf32(v1, h1, l1);
      if (vec) {
        *reinterpret_cast<float2*>(e.out + o) = make_float2(h0, h1);
        *reinterpret_cast<float2*>(e.out_lo + o) = make_float2(l0, l1);
      } else {
        e.out[o] = h0; e.out_lo[o] = l0;
        if (has1) { e.out[o + 1] = h1; e.out_lo[o + 1] = l1; }
      }
    } else if (e.out) {
      if (vec) *reinterpret_cast<float2*>(e.out + o) = make_float2(v0, v1);
      else { e.out[o] = v0; if (has1) e.out[o + 1] = v1; }
    }
    if (e.out_bf16) {
      if (vec) *reinterpret_cast<__nv_bfloat162*>(e.out_bf16 + o) = __floats2bfloat162_rn(v0, v1);
      else { e.out_bf16[o] = __float2bfloat16_rn(v0); if (has1) e.out_bf16[o + 1] = __float2bfloat16_rn(v1); }
    }
  } else if (epi == EPI_RESID) {
    size_t o = (size_t)m * e.ld_out + col;
    e.out[o] = e.resid[o] + v0;
    if (has1) e.out[o + 1] = e.resid[o + 1] + v1;
  } else if (epi == EPI_SWIGLU) {
    float h = silu_ref(v0) * v1;
    size_t o = (size_t)m * e.ld_out + (col >> 1);
    if (e.out_lo) {
      float hh, hl;
      split_tf32(h, hh, hl);
      e.out[o] = hh;
      e.out_lo[o] = hl;
    } else if (e.out) e.out[o] = h;
    if (e.out_bf16) e.out_bf16[o] = __float2bfloat16_rn(h);
  } else if (epi == EPI_ROPE_KV) {
    const int b = m / e.L, t = m - b * e.L;
    const int pos = (e.row_pos ? e.row_pos[b] : *e.pos_ptr) + t;
    const bool real = !e.row_len || t < e.row_len[b];
    const int qcols = e.HN * e.HD, kcols = e.KVHN * e.HD;
    if (col < qcols + kcols) {
      const int within = (col < qcols ? col : col - qcols);
      const int j = (within % e.HD) >> 1;
      const float c = e.cos_tab[(size_t)pos * (e.HD >> 1) + j];
      const float s = e.sin_tab[(size_t)pos * (e.HD >> 1) + j];
      const float r0 = v0 * c - v1 * s;
      const float r1 = v0 * s + v1 * c;
      if (col < qcols) {
        size_t o = (size_t)m * e.ld_out + col;
        if (e.out) { e.out[o] = r0; e.out[o + 1] = r1; }
        if (e.out_bf16) { e.out_bf16[o] = __float2bfloat16_rn(r0); e.out_bf16[o + 1] = __float2bfloat16_rn(r1); }
      } else if (real) {
        const int h = within / e.HD, d = within % e.HD;
        KVT* ck = (KVT*)e.cache_k + (((size_t)b * e.KVHN + h) * e.M + pos) * e.HD + d;
        ck[0] = from_f32<KVT>(r0);
        ck[1] = from_f32<KVT>(r1);
      }
    } else if (real) {
      const int within = col - qcols - kcols;
      const int h = within / e.HD, d = within % e.HD;
      KVT* cv = (KVT*)e.cache_v + (((size_t)b * e.KVHN + h) * e.M + pos) * e.HD + d;
      cv[0] = from_f32<KVT>(v0);
      cv[1] = from_f32<KVT>(v1);
    }
  }
}

// ------------------------------------------------------------------ launch interface
struct LinearArgs {
  const void* W;        // [N, K] of WT, row-major (the reference's [out, in] layout)
  const float* x;       // [*, K] fp32 activations
  int rows;             // activation rows
  int N, K;
  // optional fused RMSNorm of the input rows (llama3.py:111-114) while staging them
  const float* norm_w;  // null = no norm
  float eps;
  int src_mul, src_add; // source row of activation row m = m * src_mul + src_add
  const int32_t* src_rows;  // ragged batches: explicit source row per activation row (null: the affine map)
  int epi;
  EpiArgs e;
  int l2_prefetch_pairs;  // GEMV: row pairs per warp requested from L2 ahead of the dependency wait
};

// all launchers return cudaGetLastError() of the launch
cudaError_t launch_linear_rows(const LinearArgs& a, bool w_bf16, bool kv_bf16, cudaStream_t s);   // GEMV family, rows <= 8
cudaError_t launch_linear_simt(const LinearArgs& a, bool w_bf16, bool kv_bf16, cudaStream_t s);   // tiled FFMA GEMM
bool linear_rows_supported(int rows, int K);

// row r = (b, t) = (r / L, r % L) reads token ids[b * ids_ld + ids_off + t]
cudaError_t launch_embed(const void* table, bool bf16_table, const int32_t* ids, int ids_ld, int ids_off, int L,
                         int rows, int D, int vocab, float* x, cudaStream_t s);
// out_lo != null: out receives the TF32 hi part and out_lo the lo part (3xTF32 GEMM operands)
cudaError_t launch_rmsnorm(const float* x, const float* w, float eps, int rows, int D, int src_mul, int src_add,
                           float* out, bf16* out_bf16, float* out_lo, cudaStream_t s, const int32_t* src_rows = nullptr);
cudaError_t launch_argmax(const float* logits, int rows, int n, int32_t* next_ids, int64_t* out64,
                          int out_stride, const int* step_ptr, cudaStream_t s);
// per-row (value, first index) keys with a global column offset (vocabulary-sharded LM head)
cudaError_t launch_argmax_keys(const float* logits, int rows, int n, int col_offset, unsigned long long* best,
                               cudaStream_t s);
cudaError_t launch_gather_permute(const float* in, int G, int rows, int n, float* out, cudaStream_t s);
// best[rows] keys -> next_ids / out64 (as launch_argmax), and resets the keys to zero
cudaError_t launch_argmax_finalize(unsigned long long* best, int rows, int32_t* next_ids, int64_t* out64,
                                   int out_stride, const int* step_ptr, cudaStream_t s);
cudaError_t launch_set_int(int* p, int v, cudaStream_t s);
cudaError_t launch_add_int(int* p, int v, cudaStream_t s);
cudaError_t launch_rope_only(const float* x, const float* cos_tab, const float* sin_tab, int B, int L, int heads,
                             int HD, const int* pos_ptr, float* out, cudaStream_t s);
cudaError_t launch_swiglu(const float* gate, const float* up, int64_t n, float* out, cudaStream_t s);
cudaError_t launch_pack_rows(const float* src, int rows, int cols, void* dst, bool dst_bf16, int dst_row0,
                             int dst_row_stride, int dst_ld, cudaStream_t s);
cudaError_t launch_ragged_setup(const int* len, int B, int Lmax, int32_t* lastrow, int* done, cudaStream_t s);
cudaError_t launch_ragged_advance(int* scal, const int* len, int off, int B, int* rowpos, cudaStream_t s);
cudaError_t launch_ragged_eos(int32_t* next_ids, int* done, int eos, int B, int64_t* tokens, int stride, const int* step_ptr,
                              cudaStream_t s);
cudaError_t launch_add_bf16(float* x, const bf16* d, int64_t n, cudaStream_t s);  // x += d, n % 8 == 0
cudaError_t launch_unpack_bf16(const bf16* in, int64_t n, float* out, cudaStream_t s);
cudaError_t launch_fill_random(void* dst, bool dst_bf16, int64_t rows, int64_t cols, int64_t ld_global,
                               int64_t row0_global, int64_t col0_global, int dst_row0, int dst_row_stride,
                               int dst_ld, uint64_t seed, uint32_t tensor_id, float scale, float bias,
                               cudaStream_t s);
cudaError_t launch_cache_to_ref_layout(const void* cache, bool kv_bf16, int maxB, int KVHN, int M, int HD,
                                       float* out, cudaStream_t s);
cudaError_t launch_cache_from_ref_layout(const float* in, bool kv_bf16, int B, int T, int KVHN, int M, int HD,
                                         void* cache, cudaStream_t s);

struct AttnArgs {
  const float* q;      // [B*L, HN*HD] fp32, already rotated
  const void* cache_k; // [maxB, KVHN, M, HD] of KVT
  const void* cache_v;
  float* out;          // [B*L, HN*HD]
  bf16* out_bf16;      // optional mirror
  float* out_lo;       // if set: out = TF32 hi part, out_lo = lo part
  const int* pos_ptr;  // device scalar start_pos; keys visible to query t: [0, start_pos + t]
  const int* row_pos;  // ragged decode: per-sequence position [B] (null: *pos_ptr for all)
  int B, L, HN, KVHN, HD, M;
  // decode split-KV scratch (L == 1): partial o [B, HN, nsplit, HD], (m, l) [B, HN, nsplit, 2]
  float* part_o;
  float* part_ml;
  int nsplit;
  int* counters;       // [B * head groups] zero-initialised arrival counters: the last split CTA merges
                       // (null: a separate attn_combine_kernel launch merges)
  int cache_rows;      // tensor-core prefill: maxB * KVHN * M rows of the cache (TMA tensor map extent)
  // decode inside the persistent kernel: the output also / instead goes out as 8-byte {value, tag} words that the
  // consuming phase polls (flag-in-data: no grid barrier between attention and the output projection)
  unsigned long long* out_ll;  // [B*L, HN*HD] words (null: off)
  unsigned out_tag;
  int force_exact;     // decode over a bf16 cache: never the tensor-core kernel (which rounds q and the softmax weights to bf16)
};
cudaError_t launch_attn_decode(const AttnArgs& a, bool kv_bf16, cudaStream_t s);   // L == 1
// whether launch_attn_decode runs the tensor-core (mma.sync) kernel for this shape once the launch has >= 96 CTAs
bool attn_decode_mma_eligible(int HD, int nrep, bool kv_bf16);
cudaError_t launch_attn_prefill(const AttnArgs& a, bool kv_bf16, cudaStream_t s);  // L > 1
bool attn_head_dim_supported(int HD);
// bf16 tensor-core prefill (attention_tc.cu): q16 [B*L, HN*HD] bf16, bf16 caches, out = a.out_bf16
bool attn_prefill_tc_supported(int HD);
cudaError_t launch_attn_prefill_tc(const AttnArgs& a, const bf16* q16, cudaStream_t s);
