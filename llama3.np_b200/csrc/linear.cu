// Projections y = x @ W^T with W stored [out, in] (the reference's `x @ self.*_weight` lines,
// llama3.py:99-102, 166-168, 211, 307), in two SIMT forms:
//
//  * linear_rows_kernel  - row-streaming GEMV for <= 8 activation rows (batch-1 / small-batch
//    decode).  HBM-bound: every weight byte is read exactly once with 16-byte coalesced
//    streaming loads, activations are staged once per CTA in shared memory - with the RMSNorm
//    of llama3.py:111-114 fused into that staging - and reduced with warp shuffles.
//  * linear_simt_kernel  - 64x64x16 register-tiled FFMA GEMM for many rows in fp32 mode
//    (fp32 accuracy is required for token identity; the bf16 tensor-core path is gemm_tc.cu).
//
// Both finish through epilogue_pair (common.cuh): plain store, residual add, SwiGLU, or
// RoPE + KV-cache append.
#include <stdlib.h>

#include "common.cuh"

// ============================================================================ GEMV family
// One warp owns a PAIR of adjacent weight rows (so RoPE pairs and gate/up pairs land in one
// thread) and loops over K in 16-byte steps per lane; MB activation rows share each weight
// load.  Work assignment is SM-balanced: the grid is a whole number of CTAs per SM and row pair
// u goes to warp (u / gridDim) of CTA (u % gridDim), so every SM streams the same number of rows
// (within one) from the first cycle to the last.  The first weight vectors are requested BEFORE
// the activation rows are staged - weights do not depend on the previous kernel - so the HBM
// stream starts while RMSNorm staging (and, under PDL, the predecessor's tail) is still running.
template <typename WT, int MB, int EPI, typename KVT, int PF>
__global__ void __launch_bounds__(512) linear_rows_kernel(LinearArgs a) {
  extern __shared__ __align__(16) float xs[];  // [MB][K]
  constexpr int VEC = Vec16<WT>::N;            // PF = weight vectors per row in flight per lane
  const int K = a.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const WT* W = reinterpret_cast<const WT*>(a.W);
  const int npairs = (a.N + 1) >> 1;
  const int stride = nwarp * gridDim.x;
  pdl_launch();

  // ---- request the head of this warp's first row pair: weights do not depend on the previous kernel
  int p = warp * gridDim.x + blockIdx.x;
  uint4 pre0[PF], pre1[PF];
  if (p < npairs) {
    const int r0 = 2 * p;
    const WT* w0 = W + (size_t)r0 * K;
    const WT* w1 = W + (size_t)((r0 + 1) < a.N ? r0 + 1 : r0) * K;
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int k = (lane + 32 * i) * VEC;
      if (k < K) { pre0[i] = ldg_stream16(w0 + k); pre1[i] = ldg_stream16(w1 + k); }
    }
  }
  // ---- ask the L2 for this warp's first row pairs as whole rows (one bulk prefetch each):
  // the HBM stream keeps running through the dependency wait and the activation staging
  if (a.l2_prefetch_pairs > 0 && lane == 0) {
    int pp = p;
    for (int i = 0; i < a.l2_prefetch_pairs && pp < npairs; ++i, pp += stride) {
      const int nrow = (2 * pp + 1 < a.N) ? 2 : 1;
      l2_prefetch_bulk(W + (size_t)2 * pp * K, (uint32_t)(nrow * K * sizeof(WT)));
    }
  }
  pdl_wait();

  // ---- stage (and optionally RMS-normalise) the activation rows: the whole CTA works on each row
  __shared__ float red_ss[32];
  for (int m = 0; m < MB; ++m) {
    float* dst = xs + (size_t)m * K;
    if (m < a.rows) {
      const float* src = a.x + (a.src_rows ? (size_t)a.src_rows[m] : (size_t)m * a.src_mul + a.src_add) * K;
      float ss = 0.f;
      for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        *reinterpret_cast<float4*>(dst + k) = v;
      }
      if (a.norm_w) {
        ss = warp_sum(ss);
        if (lane == 0) red_ss[warp] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < nwarp; ++w) tot += red_ss[w];  // same order in every thread
        const float rinv = 1.0f / sqrtf(tot / (float)K + a.eps);
        for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4) {  // each thread rescales what it stored
          float4 v = *reinterpret_cast<const float4*>(dst + k);
          float4 g = *reinterpret_cast<const float4*>(a.norm_w + k);
          v.x = v.x * rinv * g.x; v.y = v.y * rinv * g.y; v.z = v.z * rinv * g.z; v.w = v.w * rinv * g.w;
          *reinterpret_cast<float4*>(dst + k) = v;
        }
        __syncthreads();  // red_ss is reused by the next row
      }
    } else {
      for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + k) = make_float4(0, 0, 0, 0);
    }
  }
  __syncthreads();

  auto fma_vec = [&](const uint4& r0v, const uint4& r1v, int k, float (&acc0)[MB], float (&acc1)[MB]) {
    float a0[VEC], a1[VEC];
    Vec16<WT>::unpack(r0v, a0);
    Vec16<WT>::unpack(r1v, a1);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const float* xr = xs + (size_t)m * K + k;
#pragma unroll
      for (int v = 0; v < VEC; v += 4) {
        float4 xv = *reinterpret_cast<const float4*>(xr + v);
        acc0[m] = fmaf(a0[v], xv.x, acc0[m]); acc1[m] = fmaf(a1[v], xv.x, acc1[m]);
        acc0[m] = fmaf(a0[v + 1], xv.y, acc0[m]); acc1[m] = fmaf(a1[v + 1], xv.y, acc1[m]);
        acc0[m] = fmaf(a0[v + 2], xv.z, acc0[m]); acc1[m] = fmaf(a1[v + 2], xv.z, acc1[m]);
        acc0[m] = fmaf(a0[v + 3], xv.w, acc0[m]); acc1[m] = fmaf(a1[v + 3], xv.w, acc1[m]);
      }
    }
  };

  bool first = true;
  for (; p < npairs; p += stride, first = false) {
    const int r0 = 2 * p;
    const bool has1 = (r0 + 1) < a.N;
    const WT* w0 = W + (size_t)r0 * K;
    const WT* w1 = W + (size_t)(has1 ? r0 + 1 : r0) * K;
    float acc0[MB], acc1[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
    int kbase = 0;
    if (first) {  // consume the prefetched head
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int k = (lane + 32 * i) * VEC;
        if (k < K) fma_vec(pre0[i], pre1[i], k, acc0, acc1);
      }
      kbase = 32 * PF * VEC;
    }
    for (int kb = kbase; kb < K; kb += 32 * PF * VEC) {
      uint4 c0[PF], c1[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int k = kb + (lane + 32 * i) * VEC;
        if (k < K) { c0[i] = ldg_stream16(w0 + k); c1[i] = ldg_stream16(w1 + k); }
      }
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int k = kb + (lane + 32 * i) * VEC;
        if (k < K) fma_vec(c0[i], c1[i], k, acc0, acc1);
      }
    }
#pragma unroll
    for (int m = 0; m < MB; ++m) { acc0[m] = warp_sum(acc0[m]); acc1[m] = warp_sum(acc1[m]); }
    // lane m finishes activation row m
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (lane == m && m < a.rows) epilogue_pair<KVT>(EPI, a.e, m, r0, acc0[m], acc1[m], has1);
  }
}

bool linear_rows_supported(int rows, int K) {
  if (rows < 1 || rows > 8) return false;
  int mb = rows <= 1 ? 1 : rows <= 2 ? 2 : rows <= 4 ? 4 : 8;
  return (size_t)mb * K * sizeof(float) <= 160 * 1024 && (K % 8) == 0;
}

static int gemv_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename WT, int MB, int EPI, typename KVT>
static cudaError_t launch_rows_t(const LinearArgs& a_in, cudaStream_t s) {
  static const int l2pf = gemv_env("L3_GEMV_L2PF", 0);
  LinearArgs a = a_in;
  a.l2_prefetch_pairs = l2pf;
  // prefetch depth: 8 vectors per row per lane for the single-row kernel (tunable: L3_GEMV_PF)
  static const int pf = gemv_env("L3_GEMV_PF", 4);
  auto kern = (MB == 1 && pf == 8) ? linear_rows_kernel<WT, MB, EPI, KVT, (MB == 1 ? 8 : 4)>
                                   : linear_rows_kernel<WT, MB, EPI, KVT, 4>;
  const size_t smem = (size_t)MB * a.K * sizeof(float);
  if (smem + 2048 > 48 * 1024) {  // the 48 KB default limit covers static + dynamic shared memory: opt in near it, too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int npairs = (a.N + 1) / 2;
  // choose warps per SM so that npairs / (148 * warps) sits just below an integer (balanced SMs)
  static const int max_wps = gemv_env("L3_GEMV_WPS", MB <= 2 ? 32 : 16);
  int best_w = 4;
  double best_eff = 0.0;
  for (int w = 4; w <= max_wps; ++w) {
    const long slots = 148L * w;
    const long rounds = (npairs + slots - 1) / slots;
    double eff = (double)npairs / (double)(rounds * slots);
    if (rounds == 1 && npairs < slots) eff = (double)npairs / (double)slots;  // not enough rows: prefer fewer warps
    eff += 1e-4 * w;                                                           // ties: more warps in flight
    if (eff > best_eff) { best_eff = eff; best_w = w; }
  }
  int ctas_per_sm = 1, warps_per_cta = best_w;
  if (best_w > 16) { ctas_per_sm = 2; warps_per_cta = (best_w + 1) / 2; }
  if (npairs < 148 * 4) {  // tiny matrices: spread 2-warp CTAs over as many SMs as there are rows
    warps_per_cta = 2;
    const int grid = (npairs + 1) / 2;
    return launch_k(kern, dim3(grid), dim3(64), smem, s, a);
  }
  return launch_k(kern, dim3(148 * ctas_per_sm), dim3(32 * warps_per_cta), smem, s, a);
}

template <typename WT, int MB, typename KVT>
static cudaError_t launch_rows_e(const LinearArgs& a, cudaStream_t s) {
  switch (a.epi) {
    case EPI_STORE: return launch_rows_t<WT, MB, EPI_STORE, KVT>(a, s);
    case EPI_RESID: return launch_rows_t<WT, MB, EPI_RESID, KVT>(a, s);
    case EPI_SWIGLU: return launch_rows_t<WT, MB, EPI_SWIGLU, KVT>(a, s);
    default: return launch_rows_t<WT, MB, EPI_ROPE_KV, KVT>(a, s);
  }
}

template <typename WT, typename KVT>
static cudaError_t launch_rows_m(const LinearArgs& a, cudaStream_t s) {
  if (a.rows <= 1) return launch_rows_e<WT, 1, KVT>(a, s);
  if (a.rows <= 2) return launch_rows_e<WT, 2, KVT>(a, s);
  if (a.rows <= 4) return launch_rows_e<WT, 4, KVT>(a, s);
  return launch_rows_e<WT, 8, KVT>(a, s);
}

cudaError_t launch_linear_rows(const LinearArgs& a, bool w_bf16, bool kv_bf16, cudaStream_t s) {
  if (w_bf16) return kv_bf16 ? launch_rows_m<bf16, bf16>(a, s) : launch_rows_m<bf16, float>(a, s);
  return kv_bf16 ? launch_rows_m<float, bf16>(a, s) : launch_rows_m<float, float>(a, s);
}

// ============================================================================ SIMT GEMM
// C[M, N] = A[M, K] * W[N, K]^T, fp32 FFMA.  64x64 tile, BK = 16, 256 threads, 4x4 outputs
// per thread, register-prefetch double buffering (one __syncthreads per k-tile).
// A rows may be RMS-normalised beforehand by rmsnorm_kernel (the GEMM path keeps the norm
// as its own pass; the row-streaming path above fuses it).
#define GT_BM 64
#define GT_BN 64
#define GT_BK 16
#define GT_LD (GT_BM + 4)

template <typename WT>
__device__ __forceinline__ void load4(const WT* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

template <typename WT, int EPI, typename KVT>
__global__ void __launch_bounds__(256) linear_simt_kernel(LinearArgs a) {
  __shared__ __align__(16) float As[2][GT_BK][GT_LD];
  __shared__ __align__(16) float Bs[2][GT_BK][GT_LD];
  const int tid = threadIdx.x;
  pdl_launch();
  pdl_wait();
  const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
  const int M = a.rows, N = a.N, K = a.K;
  const WT* W = reinterpret_cast<const WT*>(a.W);

  // global -> register staging: each thread moves one 4-wide K slice of one A row and one W row
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int arow = m0 + lrow, brow = n0 + lrow;
  const int arow_c = arow < M ? arow : 0;
  const float* aptr = a.x + (a.src_rows ? (size_t)a.src_rows[arow_c] : (size_t)arow_c * a.src_mul + a.src_add) * K;
  const WT* bptr = W + (size_t)(brow < N ? brow : 0) * K;
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
    const int k = k0 + lk;
    if (arow < M && k < K) load4<float>(aptr + k, ra); else { ra[0] = ra[1] = ra[2] = ra[3] = 0.f; }
    if (brow < N && k < K) load4<WT>(bptr + k, rb); else { rb[0] = rb[1] = rb[2] = rb[3] = 0.f; }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[buf][lk + i][lrow] = ra[i]; Bs[buf][lk + i][lrow] = rb[i]; }
  };

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0);
  __syncthreads();
  const int nk = (K + GT_BK - 1) / GT_BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * GT_BK);
#pragma unroll
    for (int k = 0; k < GT_BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j += 2) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      epilogue_pair<KVT>(EPI, a.e, m, col, acc[i][j], acc[i][j + 1], col + 1 < N);
    }
  }
}

template <typename WT, typename KVT>
static cudaError_t launch_simt_e(const LinearArgs& a, cudaStream_t s) {
  dim3 grid((a.N + GT_BN - 1) / GT_BN, (a.rows + GT_BM - 1) / GT_BM);
  switch (a.epi) {
    case EPI_STORE: return launch_k(linear_simt_kernel<WT, EPI_STORE, KVT>, grid, dim3(256), 0, s, a);
    case EPI_RESID: return launch_k(linear_simt_kernel<WT, EPI_RESID, KVT>, grid, dim3(256), 0, s, a);
    case EPI_SWIGLU: return launch_k(linear_simt_kernel<WT, EPI_SWIGLU, KVT>, grid, dim3(256), 0, s, a);
    default: return launch_k(linear_simt_kernel<WT, EPI_ROPE_KV, KVT>, grid, dim3(256), 0, s, a);
  }
}

cudaError_t launch_linear_simt(const LinearArgs& a, bool w_bf16, bool kv_bf16, cudaStream_t s) {
  if (w_bf16) return kv_bf16 ? launch_simt_e<bf16, bf16>(a, s) : launch_simt_e<bf16, float>(a, s);
  return kv_bf16 ? launch_simt_e<float, bf16>(a, s) : launch_simt_e<float, float>(a, s);
}
