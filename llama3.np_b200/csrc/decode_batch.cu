// Batched decode step (many sequences, one new token each: Llama.generate's per-token forward,
// llama3.py:285-321 with L == 1 and B > 128) as ONE persistent kernel, fp32 mode.
//
// Why: at stories15M shape the step is ~47 small kernels of 8-15 us, each of which pays a grid launch, a
// prologue (barrier init, TMEM allocation, tensor-map fetch), a pipeline fill and a drain for ~5 us of
// useful work - and programmatic dependent launch cannot hide that, because every kernel still waits
// for the complete predecessor.  Here one CTA per SM keeps its TMEM allocation, its TMA ring and its
// mbarriers for the whole step and walks the phases
//     [embed +] RMSNorm | QKV GEMM (RoPE + KV append) | attention | Wo GEMM (+residual) | RMSNorm |
//     W1/W3 GEMM (SwiGLU) | W2 GEMM (+residual)          per layer, then RMSNorm | LM head (argmax)
// separated by grid barriers (~1 us) instead of kernel boundaries.  The GEMM phases are the very
// pipeline of gemm_tc.cu (tc_gemm_run: TMA producer warp, tcgen05 issuer warp, four epilogue warps),
// the attention phase is attn_decode_item with one warp per (sequence, head group).
// Activations between phases live in L2; phases read them through TMA (bypasses L1) or ld.global.cg.
#include <stdio.h>

#include "attn_decode.cuh"
#include "batch.h"
#include "gemm_tc_dev.cuh"

namespace {
constexpr int BT_KIND = TC_TF32X3, BT_BN = 32, BT_THREADS = 192;
using BtCfg = TcCfg<BT_KIND, BT_BN>;

__device__ __forceinline__ unsigned bt_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid barrier over all threads of all CTAs (monotonic arrival counter, see decode_mega.cu)
__device__ __forceinline__ void bt_grid_sync(const BatchArgs& a, uint32_t& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar_cnt) : "memory");
    uint32_t spins = 0;
    while ((int32_t)(bt_ld_acquire(a.bar_cnt) - target) < 0)
      if (++spins > (1u << 26)) __trap();
    // activations written by other SMs through the generic proxy are read next through TMA
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  target += gridDim.x;
  __syncthreads();
}

// xn (hi, lo) = RMSNorm(src row) (llama3.py:111-114), ONE WARP PER ROW: 6 warps x 148 CTAs cover 888 rows in
// a single pass, the row stays in registers between the two passes (D <= 1024), nothing is synchronised.
// Layer 0 reads the embedding row of the sequence's previous token and also publishes the residual stream x.
constexpr int BT_NV = 8;  // float4s per lane kept in registers: D <= 32 * 4 * 8 = 1024
__device__ __forceinline__ void bt_norm_rows(const BatchArgs& a, const float* w, bool from_embed, float* out_hi, float* out_lo,
                                             float* /*red*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, D = a.D;
  for (int r = blockIdx.x * 6 + warp; r < a.B; r += gridDim.x * 6) {
    const float* src = from_embed ? (const float*)a.embed + (size_t)a.d_next[r] * D : a.x + (size_t)r * D;
    float4 v[BT_NV];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < BT_NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < D) {
        v[i] = from_embed ? *reinterpret_cast<const float4*>(src + k) : __ldcg(reinterpret_cast<const float4*>(src + k));
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    }
    for (int k = (lane + 32 * BT_NV) * 4; k < D; k += 128) {  // D > 1024: the tail is read twice
      const float4 t = from_embed ? *reinterpret_cast<const float4*>(src + k) : __ldcg(reinterpret_cast<const float4*>(src + k));
      ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
    ss = warp_sum(ss);
    const float rinv = 1.0f / sqrtf(ss / (float)D + a.eps);
    auto emit = [&](int k, const float4& x4) {
      const float4 g = *reinterpret_cast<const float4*>(w + k);
      if (from_embed) __stcg(reinterpret_cast<float4*>(a.x + (size_t)r * D + k), x4);
      float4 n = make_float4(x4.x * rinv * g.x, x4.y * rinv * g.y, x4.z * rinv * g.z, x4.w * rinv * g.w), hi, lo;
      split_tf32(n.x, hi.x, lo.x); split_tf32(n.y, hi.y, lo.y); split_tf32(n.z, hi.z, lo.z); split_tf32(n.w, hi.w, lo.w);
      __stcg(reinterpret_cast<float4*>(out_hi + (size_t)r * D + k), hi);
      __stcg(reinterpret_cast<float4*>(out_lo + (size_t)r * D + k), lo);
    };
#pragma unroll
    for (int i = 0; i < BT_NV; ++i) {
      const int k = (lane + 32 * i) * 4;
      if (k < D) emit(k, v[i]);
    }
    for (int k = (lane + 32 * BT_NV) * 4; k < D; k += 128)
      emit(k, from_embed ? *reinterpret_cast<const float4*>(src + k) : __ldcg(reinterpret_cast<const float4*>(src + k)));
  }
}

template <int EPI>
__device__ __forceinline__ void bt_gemm(TcPipe& p, const BatchGemm& g) {
  tc_gemm_run<BT_KIND, BT_BN, EPI>(p, &g.maps[0], &g.maps[1], &g.maps[2], &g.maps[3], g.rows, g.N, g.K, g.a_box, 1, nullptr,
                                   nullptr, g.e, blockIdx.x, gridDim.x);
}

struct BtWarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

template <int HD, int NREP>
__global__ void __launch_bounds__(BT_THREADS, 1) decode_batch_kernel(const __grid_constant__ BatchArgs a) {
  extern __shared__ uint8_t smem_raw[];
  TcPipe p;
  tc_pipe_setup<BT_KIND, BT_BN>(p, smem_raw, a.nst);
  // scratch of the non-GEMM phases: the epilogue staging area (idle outside the GEMM phases)
  uint8_t* scratch = smem_raw + (p.epi0 - p.raw);
  float* red = reinterpret_cast<float*>(scratch);
  using ASm = AttnDecodeSmem<HD, NREP, 1, float>;
  static_assert(6 * sizeof(ASm) + 64 <= BtCfg::EPI_BYTES, "attention scratch fits the epilogue staging area");
  ASm* asm_ = reinterpret_cast<ASm*>(scratch + 64);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int step = a.scal[1] + 1;            // llama3.py:316-318: decode step i runs at pos = base + i
  const int pos = a.scal[2] + step;
  if (blockIdx.x == 0 && tid == 0) a.scal[0] = pos;  // the epilogues / attention read the position here (after a barrier)
  uint32_t target = *reinterpret_cast<volatile unsigned*>(a.bar_gen) + gridDim.x;

  for (int l = 0; l < a.NL; ++l) {
    const BatchLayer& ly = a.layers[l];
    bt_norm_rows(a, ly.norm_in, l == 0, a.xn, a.xn_lo, red);                   // llama3.py:287, 248
    bt_grid_sync(a, target);
    bt_gemm<EPI_ROPE_KV>(p, ly.qkv);                                            // llama3.py:166-187
    bt_grid_sync(a, target);
    {                                                                           // llama3.py:190-207
      AttnArgs at = a.attn;
      at.cache_k = ly.ck; at.cache_v = ly.cv;
      const int ngrp = a.HN / NREP, nitems = a.B * ngrp;
      for (int item = blockIdx.x * 6 + warp; item < nitems; item += gridDim.x * 6)
        attn_decode_item<HD, NREP, float, 1, true, BtWarpSync, 4>(at, a.HN / a.KVHN, 0, item % ngrp, ngrp, item / ngrp, pos + 1,
                                                                tid & 31, asm_[warp], BtWarpSync());
    }
    bt_grid_sync(a, target);
    bt_gemm<EPI_RESID>(p, ly.wo);                                               // llama3.py:210-211, 253
    bt_grid_sync(a, target);
    bt_norm_rows(a, ly.norm_post, false, a.xn, a.xn_lo, red);                   // llama3.py:256
    bt_grid_sync(a, target);
    bt_gemm<EPI_SWIGLU>(p, ly.w13);                                             // llama3.py:99-101
    bt_grid_sync(a, target);
    bt_gemm<EPI_RESID>(p, ly.w2);                                               // llama3.py:102, 259
    bt_grid_sync(a, target);
  }
  bt_norm_rows(a, a.norm_final, false, a.xlast, a.xlast_lo, red);               // llama3.py:304
  bt_grid_sync(a, target);
  bt_gemm<EPI_ARGMAX>(p, *a.lm);                                                // llama3.py:307, 320
  bt_grid_sync(a, target);
  if (blockIdx.x == 0) {
    for (int b = tid; b < a.B; b += BT_THREADS) {
      const unsigned long long k = __ldcg(a.d_best + b);
      a.d_best[b] = 0ull;
      const int idx = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
      a.d_next[b] = idx;
      a.d_tokens[(size_t)b * a.M + step] = (int64_t)idx;
    }
    if (tid == 0) {
      a.scal[1] = step;
      *a.bar_gen = target - gridDim.x;  // = the counter now: base of the next launch
    }
  }
  tc_pipe_teardown<BT_KIND, BT_BN>(p);
}

template <int HD, int NREP>
cudaError_t launch_t(const BatchArgs& a, int grid, cudaStream_t s) {
  auto kern = decode_batch_kernel<HD, NREP>;
  const size_t smem = (size_t)a.nst * BtCfg::STAGE_BYTES + BtCfg::EPI_BYTES + 1024 + 512;
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BtCfg::SMEM);
    if (e != cudaSuccess) return e;
    done[dev & 15] = true;
  }
  kern<<<grid, BT_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}
}  // namespace

int decode_batch_stage_bytes() { return BtCfg::STAGE_BYTES; }
int decode_batch_max_stages() { return BtCfg::STAGES; }
int decode_batch_bn() { return BT_BN; }

bool decode_batch_supported(int HD, int nrep) {
  return (HD == 48 && nrep == 1) || (HD == 64 && nrep == 4) || (HD == 128 && nrep == 4);
}

cudaError_t launch_decode_batch(const BatchArgs& a, int grid, cudaStream_t s) {
  const int nrep = a.HN / a.KVHN;
  if (a.HD == 48 && nrep == 1) return launch_t<48, 1>(a, grid, s);
  if (a.HD == 64 && nrep == 4) return launch_t<64, 4>(a, grid, s);
  if (a.HD == 128 && nrep == 4) return launch_t<128, 4>(a, grid, s);
  return cudaErrorInvalidValue;
}
