// Causal GQA attention over the KV cache: softmax(q k^T / sqrt(HD) + mask) v
// (llama3.py:190-207 with softmax llama3.py:22-24 and the mask of llama3.py:293-297).
// GQA (repeat_kv, llama3.py:79-83) is index math: q head h reads kv head h / n_rep.
// The mask is never materialised: query t of the call sees keys [0, start_pos + t].
//
//  * attn_decode_kernel (L == 1): HBM-bound streaming kernel.  One CTA per (split, kv head,
//    sequence); each group of LPK lanes owns a key at a time, holding EPL = HD / LPK
//    dimensions, and keeps its own online-softmax state (m, l, o) for the NREP query heads of
//    the group, so K and V are read exactly once for all heads sharing them.  Groups are
//    merged through shared memory; with nsplit > 1 CTAs emit (m, l, o) partials that
//    attn_combine_kernel merges (flash-decoding).
//  * attn_prefill_kernel (L > 1): shared-memory tiled flash attention in fp32 FFMA,
//    16 queries x 64 keys per tile, online softmax.  (The bf16 tensor-core prefill is
//    attention_tc.cu.)
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "attn_decode.cuh"
#include "common.cuh"
#include "gemm_tc.h"

bool attn_decode_mma_eligible(int HD, int nrep, bool kv_bf16) {
  static const bool mma_on = !(getenv("L3_ATTN_MMA") && atoi(getenv("L3_ATTN_MMA")) == 0);
  return mma_on && kv_bf16 && (HD == 64 || HD == 128) && (nrep == 4 || nrep == 8);
}

bool attn_head_dim_supported(int HD) {
  return HD == 16 || HD == 32 || HD == 48 || HD == 64 || HD == 96 || HD == 128;
}

// ============================================================================ decode (L == 1)
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };

// One query head per CTA (MHA, stories15M): cap the registers so that six CTAs fit an SM - the
// ncu capture at B = 256 showed 96 registers -> 5 CTAs -> 2.08 waves, i.e. a third pass for 8 % of the work.
#ifndef L3_ATTN_MINB
#define L3_ATTN_MINB 4   // resident CTAs per SM asked of the GQA instantiations (<= 128 registers)
#endif
#ifndef L3_ATTN_U
#define L3_ATTN_U 2      // key batches in flight per lane group
#endif
template <int HD, int NREP, typename KVT>
__global__ void __launch_bounds__(128, NREP == 1 ? 6 : L3_ATTN_MINB) attn_decode_kernel(AttnArgs a, int nrep_actual) {
  __shared__ AttnDecodeSmem<HD, NREP, 4, KVT> sm;
  pdl_launch();
  pdl_wait();
  attn_decode_item<HD, NREP, KVT, 4, false, CtaSync, L3_ATTN_U>(a, nrep_actual, blockIdx.x, blockIdx.y, gridDim.y, blockIdx.z,
                                                                 (a.row_pos ? a.row_pos[blockIdx.z] : *a.pos_ptr) + 1, threadIdx.x, sm,
                                                                 CtaSync());
}

// Plenty of independent (sequence, head group) items and no key split (batched decode of many
// sequences): ONE WARP per item, four items per CTA.  Every item is resident at once (a single wave
// instead of 2 at B = 256 x 6 heads), nothing is merged through shared memory, and each warp keeps
// 4 key batches in flight to cover its serial DRAM round trips.
struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

template <int HD, int NREP, typename KVT>
__global__ void __launch_bounds__(128) attn_decode_warp_kernel(AttnArgs a, int nrep_actual, int nitems, int ngrp) {
  __shared__ AttnDecodeSmem<HD, NREP, 1, KVT> sm[4];
  pdl_launch();
  pdl_wait();
  const int warp = threadIdx.x >> 5;
  const int item = blockIdx.x * 4 + warp;
  if (item >= nitems) return;
  const int b = item / ngrp;
  constexpr int WARP_U = 4;  // key batches in flight per warp
  attn_decode_item<HD, NREP, KVT, 1, false, WarpSync, WARP_U>(a, nrep_actual, 0, item % ngrp, ngrp, b,
                                                         (a.row_pos ? a.row_pos[b] : *a.pos_ptr) + 1, threadIdx.x & 31, sm[warp],
                                                         WarpSync());
}

// Batched decode with a key split (a few hundred CTAs of a few dozen keys each): the register-streaming item above is a
// chain of dependent DRAM round trips per warp (two key batches in flight, then math, then the next two).  Here ONE
// thread requests the CTA's whole key range up front - the rows of a (sequence, kv head) are contiguous in the
// head-major cache, so a stage of SK keys is one cp.async.bulk for K and one for V - into NST shared-memory stages
// guarded by mbarriers; the warps then run the same lane-group math on staged rows (LDS.128, conflict-free: a lane
// group reads one contiguous row).  Ranges longer than NST * SK keys refill stages as they are consumed.
__device__ __forceinline__ uint32_t attn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void attn_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void attn_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // a protocol bug must fail the launch, not hang the GPU
  } while (!ok);
}

constexpr int ATTN_STAGED_MAX_NST = 8;
template <int HD, int NREP, typename KVT>
__global__ void __launch_bounds__(128, 4) attn_decode_staged_kernel(AttnArgs a, int nrep_actual, int SK, int NST) {
  using C = DecodeCfg<HD, KVT>;
  constexpr int LPK = C::LPK, EPL = C::EPL, KPW = C::KPW, NW = 4, U = 2, KSTRIDE = NW * KPW;
  extern __shared__ __align__(128) uint8_t stage_raw[];
  __shared__ AttnDecodeSmem<HD, NREP, NW, KVT> sm;
  __shared__ __align__(8) unsigned long long bars[ATTN_STAGED_MAX_NST];
  pdl_launch();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPK, sl = lane % LPK;
  const int split = blockIdx.x, grp = blockIdx.y, ngrp = gridDim.y, b = blockIdx.z;
  const int head0 = grp * NREP, kvh = head0 / nrep_actual;
  if (tid == 0) {
    for (int i = 0; i < NST; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(attn_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();  // q and this step's cache row come from the previous kernel
  const int T = (a.row_pos ? a.row_pos[b] : *a.pos_ptr) + 1;
  const int chunk = (T + a.nsplit - 1) / a.nsplit;
  const int t0 = split * chunk, t1 = min(T, t0 + chunk);
  const int nkeys = max(0, t1 - t0), nstages = (nkeys + SK - 1) / SK;
  const size_t row_bytes = (size_t)HD * sizeof(KVT);
  const uint32_t stage_bytes = (uint32_t)(SK * row_bytes);  // of K; V follows
  const KVT* kbase = (const KVT*)a.cache_k + (((size_t)b * a.KVHN + kvh) * a.M + t0) * HD;
  const KVT* vbase = (const KVT*)a.cache_v + (((size_t)b * a.KVHN + kvh) * a.M + t0) * HD;
  auto request = [&](int i) {  // stage i of the range -> buffer i % NST
    const int nk = min(SK, nkeys - i * SK);
    const uint32_t bytes = (uint32_t)(nk * row_bytes), bar = attn_smem_u32(&bars[i % NST]);
    const uint32_t dst = attn_smem_u32(stage_raw) + (uint32_t)(i % NST) * 2u * stage_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
    attn_bulk_g2s(dst, kbase + (size_t)i * SK * HD, bytes, bar);
    attn_bulk_g2s(dst + stage_bytes, vbase + (size_t)i * SK * HD, bytes, bar);
  };
  if (tid == 0)
    for (int i = 0; i < min(nstages, NST); ++i) request(i);

  const float scale = 1.0f / sqrtf((float)HD);
  float q[NREP][EPL], o[NREP][EPL], m[NREP], l[NREP];
  attn_decode_init<HD, NREP, KVT>(a, b, head0, sl, q, o, m, l);
  for (int i = 0; i < nstages; ++i) {
    attn_mbar_wait(attn_smem_u32(&bars[i % NST]), (uint32_t)(i / NST) & 1u);
    const KVT* ks = reinterpret_cast<const KVT*>(stage_raw + (size_t)(i % NST) * 2 * stage_bytes);
    const KVT* vs = reinterpret_cast<const KVT*>(stage_raw + (size_t)(i % NST) * 2 * stage_bytes + stage_bytes);
    const int nk = min(SK, nkeys - i * SK);
    for (int base = warp * KPW; base < nk; base += KSTRIDE * U) {
      float kk[U][EPL], vv[U][EPL];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = base + sub + u * KSTRIDE;
        ok[u] = t < nk;
        const int tc = ok[u] ? t : 0;  // clamped: the row's values are not used (weight exp(-inf) = 0 needs finite v)
        load_row<HD, KV_STAGED>(ks + (size_t)tc * HD, sl, kk[u]);
        load_row<HD, KV_STAGED>(vs + (size_t)tc * HD, sl, vv[u]);
      }
      attn_decode_batch<NREP, EPL, LPK, U>(q, kk, vv, ok, scale, o, m, l);
    }
    if (i + NST < nstages) {  // CTA-uniform: hand the buffer back for the stage NST ahead
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async refill
      __syncthreads();
      if (tid == 0) request(i + NST);
    }
  }
  attn_decode_finish<HD, NREP, KVT, NW>(a, split, grp, ngrp, b, tid, o, m, l, sm, CtaSync());
}

// ---------------------------------------------------------------------------- bf16 GQA decode on mma.sync
// The lane-group kernels above spend ~60 issue slots per key on a 4-head GQA group (dot products, 16-lane shuffle
// reductions, exponentials repeated by every lane of a group): at 8B batch 32 the launch is issue-bound, not
// latency-bound.  Here the group's query heads are the (zero-padded) 16 rows of an m16n8k16 tile: S = Q K^T and
// O += P V run on the tensor cores, K and V staged by TMA in the 128-byte-swizzled layout that ldmatrix reads
// without bank conflicts (V through ldmatrix.trans - no transposed copy), online softmax on the accumulator
// fragments.  Two (or four) warps per CTA, one 16-key tile each per stage.  Q and P are rounded to bf16 (the
// tcgen05 prefill does the same): bf16 mode only, fp32 mode keeps the exact lane-group kernels.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// rows 8..15 of the A operand are the zero padding of the head group: a1 = a3 = 0
__device__ __forceinline__ void mma_16816_toprows(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void attn_tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// NW warps per CTA; a stage holds one 16-key tile per warp
template <int HD, int NREP, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 2 ? 6 : 3) attn_decode_mma_kernel(const __grid_constant__ CUtensorMap tmK,
                                                                                    const __grid_constant__ CUtensorMap tmV, AttnArgs a, int NST) {
  constexpr int SK = NW * 16, NBOX = HD / 64, BOXB = SK * 128, KBYTES = NBOX * BOXB, NT = HD / 8, KS = HD / 16;
  static_assert(HD % 64 == 0 && NREP <= 8, "head_dim in 64-column boxes, the head group in the top 8 rows of the tile");
  extern __shared__ __align__(1024) uint8_t mma_stage_raw[];
  __shared__ AttnDecodeSmem<HD, NREP, NW, bf16> sm;
  __shared__ __align__(8) unsigned long long bars[ATTN_STAGED_MAX_NST];
  pdl_launch();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;  // fragment coordinates: row (head of the group), column pair
  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int head0 = kvh * NREP;
  // 128-byte swizzle needs 1024-byte aligned stages
  const uint32_t stage0 = (attn_smem_u32(mma_stage_raw) + 1023u) & ~1023u;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV));
    for (int i = 0; i < NST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(attn_smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();  // q and this step's cache row come from the previous kernel
  const int T = (a.row_pos ? a.row_pos[b] : *a.pos_ptr) + 1;
  const int chunk = (T + a.nsplit - 1) / a.nsplit;
  const int t0 = split * chunk, t1 = min(T, t0 + chunk);
  const int nkeys = max(0, t1 - t0), nstages = (nkeys + SK - 1) / SK;
  const int row0 = (b * a.KVHN + kvh) * a.M + t0;  // first cache row of the range in the [maxB * KVHN * M, HD] view
  auto request = [&](int i) {  // stage i of the range -> buffer i % NST; a box always carries SK rows (rows past the
    const uint32_t bar = attn_smem_u32(&bars[i % NST]);  // range are other cache rows or zero fill: masked below)
    const uint32_t dst = stage0 + (uint32_t)(i % NST) * 2u * KBYTES;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * KBYTES) : "memory");
#pragma unroll
    for (int x = 0; x < NBOX; ++x) {
      attn_tma_2d(dst + x * BOXB, &tmK, x * 64, row0 + i * SK, bar);
      attn_tma_2d(dst + KBYTES + x * BOXB, &tmV, x * 64, row0 + i * SK, bar);
    }
  };
  if (tid == 0)
    for (int i = 0; i < min(nstages, NST); ++i) request(i);

  // A fragments of Q: row g = head head0 + g (zero for the padding rows), k-step ks covers dimensions 16 ks .. 16 ks + 15
  uint32_t qa[KS][2];
  {
    const float* qp = a.q + ((size_t)b * a.HN + head0 + min(g, NREP - 1)) * HD;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float2 lo = __ldcg(reinterpret_cast<const float2*>(qp + ks * 16 + 2 * t));
      const float2 hi = __ldcg(reinterpret_cast<const float2*>(qp + ks * 16 + 8 + 2 * t));
      qa[ks][0] = g < NREP ? pack_bf16x2(lo.x, lo.y) : 0u;
      qa[ks][1] = g < NREP ? pack_bf16x2(hi.x, hi.y) : 0u;
    }
  }
  const float scale = 1.0f / sqrtf((float)HD);
  float o[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) { o[n][0] = 0.f; o[n][1] = 0.f; o[n][2] = 0.f; o[n][3] = 0.f; }
  float m = -INFINITY, l = 0.f;  // of row g; l is this lane's share of the row sum until the end

  const int mi = lane >> 3, mr = lane & 7;  // ldmatrix: this lane addresses row mr of matrix mi
  for (int i = 0; i < nstages; ++i) {
    attn_mbar_wait(attn_smem_u32(&bars[i % NST]), (uint32_t)(i / NST) & 1u);
    const uint32_t ksm = stage0 + (uint32_t)(i % NST) * 2u * KBYTES, vsm = ksm + KBYTES;
    const int kv = min(SK, nkeys - i * SK) - warp * 16;  // valid keys of this warp's tile
    if (kv > 0) {
      // ---- S = Q K^T: matrices (keys 0-7 | 8-15) x (dims lo | hi) of the k-step, B fragments of two 8-key tiles
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      {
        const int key = warp * 16 + (mi >> 1) * 8 + mr;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int dd = ks * 16 + (mi & 1) * 8;
          uint32_t kb[4];
          ldsm_x4(ksm + (dd >> 6) * BOXB + key * 128 + ((((dd & 63) >> 3) ^ (key & 7)) << 4), kb);
          mma_16816_toprows(s0, qa[ks][0], qa[ks][1], kb[0], kb[1]);
          mma_16816_toprows(s1, qa[ks][0], qa[ks][1], kb[2], kb[3]);
        }
      }
      // ---- online softmax of row g: this lane holds keys 2t, 2t + 1 (s0) and 8 + 2t, 8 + 2t + 1 (s1)
      const float v00 = 2 * t < kv ? s0[0] * scale : -INFINITY, v01 = 2 * t + 1 < kv ? s0[1] * scale : -INFINITY;
      const float v10 = 8 + 2 * t < kv ? s1[0] * scale : -INFINITY, v11 = 9 + 2 * t < kv ? s1[1] * scale : -INFINITY;
      float mx = fmaxf(fmaxf(v00, v01), fmaxf(v10, v11));
      mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 2));
      const float mn = fmaxf(m, mx);  // finite: the tile has at least one valid key
      const float alpha = expf(m - mn);  // m = -inf -> 0
      const float p00 = expf(v00 - mn), p01 = expf(v01 - mn), p10 = expf(v10 - mn), p11 = expf(v11 - mn);
      l = l * alpha + (p00 + p01) + (p10 + p11);
      m = mn;
#pragma unroll
      for (int n = 0; n < NT; ++n) { o[n][0] *= alpha; o[n][1] *= alpha; }
      const uint32_t pa0 = pack_bf16x2(p00, p01), pa2 = pack_bf16x2(p10, p11);
      // ---- O += P V: matrices (keys 0-7 | 8-15) x (dims d0 .. d0 + 7 | d0 + 8 .. d0 + 15), transposed on the way in
      {
        const int key = warp * 16 + (mi & 1) * 8 + mr;
#pragma unroll
        for (int np = 0; np < KS; ++np) {
          const int dd = np * 16 + (mi >> 1) * 8;
          uint32_t vb[4];
          ldsm_x4_trans(vsm + (dd >> 6) * BOXB + key * 128 + ((((dd & 63) >> 3) ^ (key & 7)) << 4), vb);
          mma_16816_toprows(o[2 * np], pa0, pa2, vb[0], vb[1]);
          mma_16816_toprows(o[2 * np + 1], pa0, pa2, vb[2], vb[3]);
        }
      }
    }
    if (i + NST < nstages) {  // CTA-uniform: hand the buffer back for the stage NST ahead
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async refill
      __syncthreads();
      if (tid == 0) request(i + NST);
    }
  }
  // ---- this warp's state of the group's heads into shared memory, then the common merge / publish / combine
  l += __shfl_xor_sync(L3_FULL, l, 1);
  l += __shfl_xor_sync(L3_FULL, l, 2);
  if (g < NREP) {
    if (t == 0) { sm.m[g][warp] = m; sm.l[g][warp] = l; }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      sm.o[g][warp][n * 8 + 2 * t] = o[n][0];
      sm.o[g][warp][n * 8 + 2 * t + 1] = o[n][1];
    }
  }
  attn_decode_finish_smem<HD, NREP, bf16, NW>(a, split, kvh, gridDim.y, b, tid, sm, CtaSync());
}

template <int HD>
__global__ void __launch_bounds__(128) attn_combine_kernel(AttnArgs a) {
  __shared__ float cmb_w[1][32];
  __shared__ float cmb_l[1];
  pdl_launch();
  pdl_wait();
  combine_splits<HD, 1, 4>(a, blockIdx.y, blockIdx.x, threadIdx.x, cmb_w, cmb_l, CtaSync());
}

template <int HD, typename KVT>
static cudaError_t launch_decode_hd(const AttnArgs& a, cudaStream_t s) {
  const int nrep = a.HN / a.KVHN;
  dim3 block(128);
  cudaError_t e;
  static const bool warp_items = !(getenv("L3_ATTN_WARP") && atoi(getenv("L3_ATTN_WARP")) == 0);
  if (warp_items && a.nsplit == 1 && (nrep == 1 || nrep == 4) && a.B * a.KVHN >= 4 * 148) {
    const int ngrp = a.KVHN, nitems = a.B * ngrp;
    if (nrep == 1) return launch_k(attn_decode_warp_kernel<HD, 1, KVT>, dim3((nitems + 3) / 4), block, 0, s, a, nrep, nitems, ngrp);
    return launch_k(attn_decode_warp_kernel<HD, 4, KVT>, dim3((nitems + 3) / 4), block, 0, s, a, nrep, nitems, ngrp);
  }
  // bf16 cache, a GQA group of 4 or 8 heads, head_dim 64 / 128, enough CTAs to fill the machine: tensor-core kernel
  if constexpr (std::is_same<KVT, bf16>::value && (HD == 64 || HD == 128)) {
    if (attn_decode_mma_eligible(HD, nrep, true) && !a.force_exact && (long long)a.nsplit * a.KVHN * a.B >= 96) {
      // four warps per CTA and no more splits than it takes to fill the machine once (pick_nsplit, l3_api.cu) - at 8B
      // batch 32, ms per decode step: 2 warps x 512 CTAs 4.56, 2 x 256 4.38, 4 x 256 4.33, 4 x 512 4.60
      static const int nw = [] { const char* v = getenv("L3_ATTN_MMA_NW"); return v && atoi(v) == 2 ? 2 : 4; }();
      const long long rows = (long long)a.B * a.KVHN * a.M;
      const CUtensorMap* tk = rows < (1ll << 31) ? tc_get_map(a.cache_k, true, (int)rows, HD, nw * 16) : nullptr;
      const CUtensorMap* tv = tk ? tc_get_map(a.cache_v, true, (int)rows, HD, nw * 16) : nullptr;
      if (tk && tv) {
        const int stage = 2 * nw * 16 * HD * 2;
        const int budget = nw == 2 ? 49152 : 65536;  // stage ring of a CTA (four warps: 64 / 96 / 128 KB measured 4.33 / 4.34 / 4.44 ms)
        const int nst = std::max(2, std::min(ATTN_STAGED_MAX_NST, budget / stage));
        const dim3 grid(a.nsplit, a.KVHN, a.B);
        auto go = [&](auto kern, bool* attr_done) {
          int dev = 0;
          cudaGetDevice(&dev);
          if (!attr_done[dev & 15]) {
            cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10);
            if (e2 != cudaSuccess) return e2;
            attr_done[dev & 15] = true;
          }
          return launch_k(kern, grid, dim3(nw * 32), (size_t)nst * stage + 1024, s, *tk, *tv, a, nst);
        };
        static bool done[4][16] = {};
        if (nw == 2) e = nrep == 4 ? go(attn_decode_mma_kernel<HD, 4, 2>, done[0]) : go(attn_decode_mma_kernel<HD, 8, 2>, done[1]);
        else e = nrep == 4 ? go(attn_decode_mma_kernel<HD, 4, 4>, done[2]) : go(attn_decode_mma_kernel<HD, 8, 4>, done[3]);
        if (e != cudaSuccess || a.nsplit == 1 || a.counters) return e;
        return launch_k(attn_combine_kernel<HD>, dim3(a.HN, a.B), dim3(128), 0, s, a);
      }
    }
  }
  // enough CTAs to fill the machine: stage each CTA's key range in shared memory with bulk copies (see the kernel)
  static const bool staged = !(getenv("L3_ATTN_STAGED") && atoi(getenv("L3_ATTN_STAGED")) == 0);
  const int nr = (nrep == 8 || nrep == 4 || nrep == 2) ? nrep : 1;
  if (staged && (long long)a.nsplit * (a.HN / nr) * a.B >= 148) {
    using C = DecodeCfg<HD, KVT>;
    constexpr int SK = 4 * C::KPW * 2;  // keys per stage = one batch of every lane group of the CTA
    constexpr int STAGE = 2 * SK * HD * (int)sizeof(KVT);
    constexpr int NST = STAGE * 2 > 40960 ? 2 : (40960 / STAGE > ATTN_STAGED_MAX_NST ? ATTN_STAGED_MAX_NST : 40960 / STAGE);
    const dim3 grid(a.nsplit, a.HN / nr, a.B);
    auto go = [&](auto kern) {
      // the four instantiations share one function-pointer type, hence one instance of this lambda: a flag per kernel
      static bool attr_done_all[4][16] = {};
      bool* attr_done = attr_done_all[nr == 8 ? 3 : nr == 4 ? 2 : nr == 2 ? 1 : 0];
      int dev = 0;
      cudaGetDevice(&dev);
      if (!attr_done[dev & 15]) {
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10);  // static + dynamic may pass 48 KB
        if (e2 != cudaSuccess) return e2;
        attr_done[dev & 15] = true;
      }
      return launch_k(kern, grid, block, (size_t)NST * STAGE, s, a, nrep, SK, NST);
    };
    if (nr == 8) e = go(attn_decode_staged_kernel<HD, 8, KVT>);
    else if (nr == 4) e = go(attn_decode_staged_kernel<HD, 4, KVT>);
    else if (nr == 2) e = go(attn_decode_staged_kernel<HD, 2, KVT>);
    else e = go(attn_decode_staged_kernel<HD, 1, KVT>);
  } else if (nrep == 8) {
    e = launch_k(attn_decode_kernel<HD, 8, KVT>, dim3(a.nsplit, a.HN / 8, a.B), block, 0, s, a, nrep);
  } else if (nrep == 4) {
    e = launch_k(attn_decode_kernel<HD, 4, KVT>, dim3(a.nsplit, a.HN / 4, a.B), block, 0, s, a, nrep);
  } else if (nrep == 2) {
    e = launch_k(attn_decode_kernel<HD, 2, KVT>, dim3(a.nsplit, a.HN / 2, a.B), block, 0, s, a, nrep);
  } else {  // n_rep 1, or an unusual ratio: one head per CTA
    e = launch_k(attn_decode_kernel<HD, 1, KVT>, dim3(a.nsplit, a.HN, a.B), block, 0, s, a, nrep);
  }
  if (e != cudaSuccess || a.nsplit == 1 || a.counters) return e;
  return launch_k(attn_combine_kernel<HD>, dim3(a.HN, a.B), dim3(128), 0, s, a);
}

template <typename KVT>
static cudaError_t launch_decode_t(const AttnArgs& a, cudaStream_t s) {
  switch (a.HD) {
    case 16: return launch_decode_hd<16, KVT>(a, s);
    case 32: return launch_decode_hd<32, KVT>(a, s);
    case 48: return launch_decode_hd<48, KVT>(a, s);
    case 64: return launch_decode_hd<64, KVT>(a, s);
    case 96: return launch_decode_hd<96, KVT>(a, s);
    case 128: return launch_decode_hd<128, KVT>(a, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_attn_decode(const AttnArgs& a, bool kv_bf16, cudaStream_t s) {
  return kv_bf16 ? launch_decode_t<bf16>(a, s) : launch_decode_t<float>(a, s);
}

// ============================================================================ prefill (L > 1)
#define PF_ROWS 16
#define PF_TK 64

template <typename KVT> __device__ __forceinline__ float4 load_kv4(const KVT* p);
template <> __device__ __forceinline__ float4 load_kv4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <> __device__ __forceinline__ float4 load_kv4<bf16>(const bf16* p) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u),
                     __uint_as_float(t.y << 16), __uint_as_float(t.y & 0xffff0000u));
}

template <int HD, typename KVT>
__global__ void __launch_bounds__(256) attn_prefill_kernel(AttnArgs a, int nrep) {
  constexpr int LDK = HD + 4, HD4 = HD / 4, LDS = PF_TK + 1;
  constexpr int NSL = (PF_ROWS * HD4 + 255) / 256;
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                       // [ROWS][HD]
  float* Ks = Qs + PF_ROWS * HD;          // [TK][LDK]
  float* Vs = Ks + PF_TK * LDK;           // [TK][LDK]
  float* Ss = Vs + PF_TK * LDK;           // [ROWS][LDS]
  float* st_m = Ss + PF_ROWS * LDS;       // [ROWS] running max
  float* st_l = st_m + PF_ROWS;           // [ROWS] running sum
  float* st_a = st_l + PF_ROWS;           // [ROWS] rescale of this tile

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / nrep;
  pdl_launch();
  pdl_wait();
  const int start = *a.pos_ptr;
  const int q0 = blockIdx.x * PF_ROWS;
  const int nq = min(PF_ROWS, a.L - q0);
  const int t_hi = start + q0 + nq;  // keys [0, t_hi) are visible to at least one query of the tile
  const float scale = 1.0f / sqrtf((float)HD);

  for (int i = tid; i < PF_ROWS * HD4; i += 256) {
    const int r = i / HD4, d4 = i % HD4;
    float4 v = make_float4(0, 0, 0, 0);
    if (r < nq) v = *reinterpret_cast<const float4*>(a.q + ((size_t)(b * a.L + q0 + r) * a.HN + head) * HD + d4 * 4);
    *reinterpret_cast<float4*>(Qs + r * HD + d4 * 4) = v;
  }
  if (tid < PF_ROWS) { st_m[tid] = -INFINITY; st_l[tid] = 0.f; st_a[tid] = 1.f; }
  float4 acc[NSL];
#pragma unroll
  for (int i = 0; i < NSL; ++i) acc[i] = make_float4(0, 0, 0, 0);

  const KVT* kbase = (const KVT*)a.cache_k + ((size_t)b * a.KVHN + kvh) * a.M * HD;
  const KVT* vbase = (const KVT*)a.cache_v + ((size_t)b * a.KVHN + kvh) * a.M * HD;

  for (int k0 = 0; k0 < t_hi; k0 += PF_TK) {
    __syncthreads();  // previous tile consumed (first pass: Q / state visible)
    for (int i = tid; i < PF_TK * HD4; i += 256) {
      const int key = i / HD4, d4 = i % HD4;
      const int t = k0 + key;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (t < t_hi) {
        kv = load_kv4<KVT>(kbase + (size_t)t * HD + d4 * 4);
        vv = load_kv4<KVT>(vbase + (size_t)t * HD + d4 * 4);
      }
      *reinterpret_cast<float4*>(Ks + key * LDK + d4 * 4) = kv;
      *reinterpret_cast<float4*>(Vs + key * LDK + d4 * 4) = vv;
    }
    __syncthreads();
    {  // scores: thread -> one key x four query rows
      const int key = tid & 63, rg = tid >> 6;
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int d4 = 0; d4 < HD4; ++d4) {
        const float4 kv = *reinterpret_cast<const float4*>(Ks + key * LDK + d4 * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + (rg * 4 + i) * HD + d4 * 4);
          s[i] = fmaf(qv.x, kv.x, s[i]); s[i] = fmaf(qv.y, kv.y, s[i]);
          s[i] = fmaf(qv.z, kv.z, s[i]); s[i] = fmaf(qv.w, kv.w, s[i]);
        }
      }
      const int t = k0 + key;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rg * 4 + i;
        const bool vis = (r < nq) && (t <= start + q0 + r);  // causal predicate (llama3.py:293-297)
        Ss[r * LDS + key] = vis ? s[i] * scale : -INFINITY;
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {  // online softmax: warp w owns rows 2w, 2w+1
      const int r = warp * 2 + rr;
      const float mold = st_m[r];
      const float s0 = Ss[r * LDS + lane], s1 = Ss[r * LDS + lane + 32];
      const float mnew = fmaxf(mold, warp_max(fmaxf(s0, s1)));
      float p0 = 0.f, p1 = 0.f, alpha = 1.f;
      if (mnew > -INFINITY) {
        p0 = expf(s0 - mnew);
        p1 = expf(s1 - mnew);
        alpha = expf(mold - mnew);
      }
      const float ps = warp_sum(p0 + p1);
      Ss[r * LDS + lane] = p0;
      Ss[r * LDS + lane + 32] = p1;
      if (lane == 0) { st_l[r] = st_l[r] * alpha + ps; st_m[r] = mnew; st_a[r] = alpha; }
    }
    __syncthreads();
#pragma unroll
    for (int si = 0; si < NSL; ++si) {  // o = o * alpha + P V
      const int sidx = tid + si * 256;
      if (sidx < PF_ROWS * HD4) {
        const int r = sidx / HD4, d4 = sidx % HD4;
        const float alpha = st_a[r];
        float4 o = acc[si];
        o.x *= alpha; o.y *= alpha; o.z *= alpha; o.w *= alpha;
#pragma unroll 8
        for (int key = 0; key < PF_TK; ++key) {
          const float p = Ss[r * LDS + key];
          const float4 v = *reinterpret_cast<const float4*>(Vs + key * LDK + d4 * 4);
          o.x = fmaf(p, v.x, o.x); o.y = fmaf(p, v.y, o.y); o.z = fmaf(p, v.z, o.z); o.w = fmaf(p, v.w, o.w);
        }
        acc[si] = o;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int si = 0; si < NSL; ++si) {
    const int sidx = tid + si * 256;
    if (sidx < PF_ROWS * HD4) {
      const int r = sidx / HD4, d4 = sidx % HD4;
      if (r < nq) {
        const float inv = 1.0f / st_l[r];
        float4 o = acc[si];
        o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        const size_t oi = ((size_t)(b * a.L + q0 + r) * a.HN + head) * HD + d4 * 4;
        if (a.out_lo) {
          float4 hi, lo;
          split_tf32(o.x, hi.x, lo.x); split_tf32(o.y, hi.y, lo.y); split_tf32(o.z, hi.z, lo.z); split_tf32(o.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(a.out + oi) = hi;
          *reinterpret_cast<float4*>(a.out_lo + oi) = lo;
        } else if (a.out) *reinterpret_cast<float4*>(a.out + oi) = o;
        if (a.out_bf16) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(a.out_bf16 + oi) = pk;
        }
      }
    }
  }
}

template <int HD, typename KVT>
static cudaError_t launch_prefill_hd(const AttnArgs& a, cudaStream_t s) {
  auto kern = attn_prefill_kernel<HD, KVT>;
  const size_t smem = (size_t)(PF_ROWS * HD + 2 * PF_TK * (HD + 4) + PF_ROWS * (PF_TK + 1) + 3 * PF_ROWS) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((a.L + PF_ROWS - 1) / PF_ROWS, a.HN, a.B);
  return launch_k(kern, grid, dim3(256), smem, s, a, a.HN / a.KVHN);
}

template <typename KVT>
static cudaError_t launch_prefill_t(const AttnArgs& a, cudaStream_t s) {
  switch (a.HD) {
    case 16: return launch_prefill_hd<16, KVT>(a, s);
    case 32: return launch_prefill_hd<32, KVT>(a, s);
    case 48: return launch_prefill_hd<48, KVT>(a, s);
    case 64: return launch_prefill_hd<64, KVT>(a, s);
    case 96: return launch_prefill_hd<96, KVT>(a, s);
    case 128: return launch_prefill_hd<128, KVT>(a, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_attn_prefill(const AttnArgs& a, bool kv_bf16, cudaStream_t s) {
  return kv_bf16 ? launch_prefill_t<bf16>(a, s) : launch_prefill_t<float>(a, s);
}
