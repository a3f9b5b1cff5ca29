// Tensor-core projections for 9..128 activation rows (batched decode): y = x W^T with the roles
// of the MMA operands SWAPPED - a tile's 128-row operand is a block of WEIGHT rows and the batch
// is the N dimension (32 / 64 / 128 columns).  With the batch as the 128-row operand
// (gemm_tc.cu) every CTA re-reads the whole activation tile to stream only 32 weight rows; here a
// CTA streams 128 weight rows per 16-128 KB of activations, so shared-memory ingress is spent on
// the bytes that matter (weights are read exactly once from HBM, the HBM-bound regime of
// llama3.py:99-102,166-168,211 at small batch).  Matrices with few row blocks (QKV: 48, Wo / Wdown:
// 32 at D = 4096) are additionally split along K: every slice publishes its partial tile and the last
// one to arrive sums them in slice order (deterministic) before the fused epilogue.  The bf16-mode
// residual epilogue adds its slices into the residual stream with fp32 atomics instead (no reduction
// tail: 4 us per GEMM faster at the 8B shape); fp32 mode never reorders a sum.
//
// The accumulator is transposed (TMEM lane = output column n, TMEM column = activation row m), so
// the epilogue needs no shared-memory pass: for a fixed m the 32 lanes of a warp hold 32
// consecutive n - every global access is coalesced as it stands; RoPE / SwiGLU pairs are adjacent
// lanes (one shuffle).
#include <cuda.h>

#include <stdlib.h>
#include <algorithm>

#include "common.cuh"
#include "gemm_tc.h"

namespace {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if constexpr (KIND == TC_BF16) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// BNA = batch columns of the tile (activation rows padded up)
template <int KIND, int BNA> struct SwCfg {
  static constexpr int PARTS = KIND == TC_TF32X3 ? 2 : 1;
  static constexpr int BMW = 128;                       // weight rows per tile
  static constexpr int BK = KIND == TC_BF16 ? 64 : 32;  // elements per 128-byte row
  static constexpr int W_BYTES = BMW * 128, X_BYTES = BNA * 128;
  static constexpr int STAGE_BYTES = PARTS * (W_BYTES + X_BYTES);
  // (bf16 with four round-robin accumulators was measured: no gain at 8B batch 32, so one accumulator)
  static constexpr int NACC = KIND == TC_TF32X3 ? 4 : 1;
  static constexpr int ACC_COLS = NACC * BNA;
  static constexpr int NBUF = 2 * ACC_COLS <= 512 ? 2 : 1;
  static constexpr int TMEM_COLS = NBUF * ACC_COLS <= 32 ? 32 : NBUF * ACC_COLS <= 64 ? 64 : NBUF * ACC_COLS <= 128 ? 128
                                   : NBUF * ACC_COLS <= 256 ? 256 : 512;
  static constexpr int STAGES_RAW = (232448 - 1024 - 512) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 10 ? 10 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 512;
  static constexpr uint32_t FMT = KIND == TC_BF16 ? 1u : 2u;
  // D fp32, both operands K-major, N = BNA, M = 128
  static constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BNA >> 3) << 17) | ((128u >> 4) << 24);
};

template <int KIND, int BNA, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_swap_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                 int rows, int N, int K, int ksplit, float* part, int* tile_cnt, EpiArgs e) {
  using Cf = SwCfg<KIND, BNA>;
  using KVT = typename std::conditional<KIND == TC_BF16, bf16, float>::type;
  constexpr int PARTS = Cf::PARTS, STAGES = Cf::STAGES, NBUF = Cf::NBUF;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t bars = tiles + STAGES * Cf::STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, accfull0 = bars + 16 * STAGES, accempty0 = accfull0 + 16,
                 tmem_slot = accempty0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (N + Cf::BMW - 1) / Cf::BMW;
  const int ntiles = tiles_w * ksplit;
  const int nkb = (K + Cf::BK - 1) / Cf::BK;
  const int kb_per = (nkb + ksplit - 1) / ksplit;
  pdl_launch();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX0));
    if (PARTS == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX1));
    }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(accfull0 + 8 * b, 1); mbar_init(accempty0 + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(Cf::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      // weights do not depend on the previous kernel: their first stages are requested before the
      // dependency wait; the activation boxes follow once the producer kernel's writes are visible
      uint32_t it = 0;
      bool waited = false;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n0 = (t % tiles_w) * Cf::BMW, ks = t / tiles_w;
        const int kb0 = ks * kb_per, kb1 = min(nkb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t st = tiles + s * Cf::STAGE_BYTES;
          mbar_expect_tx(full0 + 8 * s, PARTS * (Cf::W_BYTES + Cf::X_BYTES));
          tma_load_2d(st, &tmW0, kb * Cf::BK, n0, full0 + 8 * s);
          if (PARTS == 2) tma_load_2d(st + Cf::W_BYTES, &tmW1, kb * Cf::BK, n0, full0 + 8 * s);
          if (!waited) { pdl_wait(); waited = true; }
          tma_load_2d(st + PARTS * Cf::W_BYTES, &tmX0, kb * Cf::BK, 0, full0 + 8 * s);
          if (PARTS == 2) tma_load_2d(st + PARTS * Cf::W_BYTES + Cf::X_BYTES, &tmX1, kb * Cf::BK, 0, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      uint32_t it = 0, ti = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
        const int ks = t / tiles_w;
        const int kb0 = ks * kb_per, kb1 = min(nkb, kb0 + kb_per);
        const uint32_t buf = ti % NBUF, use = ti / NBUF;
        if (use > 0) {
          mbar_wait(accempty0 + 8 * buf, (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + buf * Cf::ACC_COLS;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = tiles + s * Cf::STAGE_BYTES;
          const uint64_t w_hi = umma_desc_sw128(st), x_hi = umma_desc_sw128(st + PARTS * Cf::W_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const int slice = (kb - kb0) * 4 + kk;
            if (PARTS == 2) {
              const uint64_t w_lo = umma_desc_sw128(st + Cf::W_BYTES);
              const uint64_t x_lo = umma_desc_sw128(st + PARTS * Cf::W_BYTES + Cf::X_BYTES);
              const uint32_t corr = acc + 3 * BNA, mainacc = acc + (slice % 3) * BNA;
              tc_mma<KIND>(corr, w_lo + adv, x_hi + adv, Cf::IDESC, slice == 0 ? 0u : 1u);
              tc_mma<KIND>(corr, w_hi + adv, x_lo + adv, Cf::IDESC, 1u);
              tc_mma<KIND>(mainacc, w_hi + adv, x_hi + adv, Cf::IDESC, slice < 3 ? 0u : 1u);
            } else {
              tc_mma<KIND>(acc + (slice % Cf::NACC) * BNA, w_hi + adv, x_hi + adv, Cf::IDESC, slice < Cf::NACC ? 0u : 1u);
            }
          }
          tc_commit(empty0 + 8 * s);
        }
        tc_commit(accfull0 + 8 * buf);
      }
    }
  } else {  // ---------------- epilogue (warps 2-5): lane = output column, TMEM column = activation row
    pdl_wait();  // residual / position reads below depend on the previous kernel
    const int quarter = warp & 3;
    const int start_pos = (EPI == EPI_ROPE_KV) ? *e.pos_ptr : 0;
    uint32_t ti = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
      const int n0 = (t % tiles_w) * Cf::BMW, ks = t / tiles_w;
      const int kb0 = ks * kb_per, kb1 = min(nkb, kb0 + kb_per);
      const uint32_t buf = ti % NBUF, use = ti / NBUF;
      mbar_wait(accfull0 + 8 * buf, use & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * Cf::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      const int n = n0 + quarter * 32 + lane;       // this lane's output column
      const bool n_ok = n < N && kb0 < kb1;
      const bool live = n0 + quarter * 32 < N && kb0 < kb1;  // warp-uniform
      auto load_acc = [&](int c0, float (&v)[32]) {
        tmem_ld32(acc + (uint32_t)c0, v);
        if (Cf::NACC == 4) {
          float w[32];
#pragma unroll
          for (int k2 = 1; k2 < 4; ++k2) {
            tmem_ld32(acc + (uint32_t)c0 + k2 * BNA, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
        }
      };
      const int Np = tiles_w * Cf::BMW;  // padded width of the K-split scratch [ksplit][BNA][Np]
      const int tw = t % tiles_w;
      const bool det = ksplit > 1 && part != nullptr;  // else (residual epilogue only): fp32 atomics
      if (det) {
        // K-split: publish this slice's partial tile (for a fixed row the warp's 32 lanes store 32 consecutive
        // columns); the LAST slice of an output tile to arrive sums all slices in slice order - a fixed order,
        // so the result does not depend on timing - and runs the fused epilogue on the sum.
#pragma unroll 1
        for (int c0 = 0; c0 < BNA; c0 += 32) {
          if (c0 >= rows || !live) break;
          float v[32];
          load_acc(c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < rows) __stcg(part + ((size_t)ks * BNA + c0 + j) * Np + n0 + quarter * 32 + lane, v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
        // release: the CTA barrier orders the four warps' partial stores before thread 64's acq_rel atomic
        // (cumulative), whose acquire side + the next barrier order the last arriver's reads after all slices'
        asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
        int* flag = reinterpret_cast<int*>(smem_raw + (tmem_slot - raw)) + 1;
        if (threadIdx.x == 64) {
          int old;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(tile_cnt + tw) : "memory");
          *flag = (old == ksplit - 1);
          if (old == ksplit - 1) tile_cnt[tw] = 0;  // ready for the next launch
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const bool is_last = *flag != 0;
        asm volatile("bar.sync 2, 128;" ::: "memory");  // everyone has read the flag before it is reused
        if (!is_last) continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BNA; c0 += 32) {
        if (c0 >= rows || !live) break;  // warp-uniform
        float v[32];
        if (!det) {
          load_acc(c0, v);
          if (c0 + 32 >= BNA || c0 + 32 >= rows) {  // last TMEM read of this tile: hand the buffer back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
          }
        } else {  // sum of the slices' partials, slice 0 first
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          // unconditional loads (the scratch holds BNA rows per slice): all 32 are in flight before the first add -
          // guarded ones are issued three at a time, ~10 dependent L2 round trips per slice
          for (int k2 = 0; k2 < ksplit; ++k2) {
            const float* src = part + ((size_t)k2 * BNA + c0) * Np + n0 + quarter * 32 + lane;
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __ldcg(src + (size_t)j * Np);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
        }
        if constexpr (EPI == EPI_STORE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c0 + j;
            if (m < rows && n_ok) {
              const size_t o = (size_t)m * e.ld_out + n;
              if (e.out_lo) { float hi, lo; split_tf32(v[j], hi, lo); e.out[o] = hi; e.out_lo[o] = lo; }
              else if (e.out) e.out[o] = v[j];
              if (e.out_bf16) e.out_bf16[o] = __float2bfloat16_rn(v[j]);
            }
          }
        } else if constexpr (EPI == EPI_RESID) {
          if (ksplit == 1 || det) {
            // clamped instead of guarded, so that all 32 loads are in flight together (values of clamped
            // duplicates are not used)
            const int nc = min(n, N - 1);
            float rr[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = e.resid[(size_t)min(c0 + j, rows - 1) * e.ld_out + nc];
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < rows && n_ok) e.out[(size_t)(c0 + j) * e.ld_out + n] = rr[j] + v[j];
          } else {  // bf16 mode: the slices add their partials into the residual stream in place (out == resid)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < rows && n_ok) atomicAdd(e.out + (size_t)(c0 + j) * e.ld_out + n, v[j]);
          }
        } else if constexpr (EPI == EPI_SWIGLU) {
          // rows interleaved gate_j, up_j: even lane holds the gate, its odd neighbour the up value
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float up = __shfl_xor_sync(L3_FULL, v[j], 1);
            const int m = c0 + j;
            if (m < rows && n_ok && !(lane & 1)) {
              const float h = silu_ref(v[j]) * up;
              const size_t o = (size_t)m * e.ld_out + (n >> 1);
              if (e.out_lo) { float hh, hl; split_tf32(h, hh, hl); e.out[o] = hh; e.out_lo[o] = hl; }
              else if (e.out) e.out[o] = h;
              if (e.out_bf16) e.out_bf16[o] = __float2bfloat16_rn(h);
            }
          }
        } else if constexpr (EPI == EPI_ROPE_KV) {
          const int qcols = e.HN * e.HD, kcols = e.KVHN * e.HD;
          const bool is_q = n < qcols, is_k = !is_q && n < qcols + kcols;
          const int within = is_q ? n : (is_k ? n - qcols : n - qcols - kcols);
          const int hh = within / e.HD, d = within % e.HD;
          // positions and table entries of all 32 rows first: as loads inside the store loop they were 32 dependent
          // L2 round trips (no load may move above a store that could alias it)
          int posv[32];
          float cs[32], sn[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = min(c0 + j, rows - 1);  // clamped: values of rows past the end are not used
            const int b = m / e.L;
            posv[j] = (e.row_pos ? __ldg(e.row_pos + b) : start_pos) + (m - b * e.L);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const size_t ti = (size_t)posv[j] * (e.HD >> 1) + (d >> 1);
            cs[j] = (is_q || is_k) ? __ldg(e.cos_tab + ti) : 1.f;
            sn[j] = (is_q || is_k) ? __ldg(e.sin_tab + ti) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float other = __shfl_xor_sync(L3_FULL, v[j], 1);  // the pair's partner (llama3.py:41-76)
            const int m = c0 + j;
            if (m >= rows || !n_ok) continue;
            const int b = m / e.L, t = m - b * e.L;
            const bool real = !e.row_len || t < e.row_len[b];
            float r = v[j];
            // even: x0 c - x1 s, odd: x0 s + x1 c
            if (is_q || is_k) r = (lane & 1) ? other * sn[j] + v[j] * cs[j] : v[j] * cs[j] - other * sn[j];
            if (is_q) {
              const size_t o = (size_t)m * e.ld_out + n;
              if (e.out) e.out[o] = r;
              if (e.out_bf16) e.out_bf16[o] = __float2bfloat16_rn(r);
            } else if (real) {
              KVT* c = (KVT*)(is_k ? e.cache_k : e.cache_v) + (((size_t)b * e.KVHN + hh) * e.M + posv[j]) * e.HD + d;
              *c = from_f32<KVT>(r);
            }
          }
        } else {  // EPI_ARGMAX (llama3.py:320): per activation row, the warp's best (value, first index)
          // two warp-wide integer reductions per row (redux.sync) on the order-preserving key halves:
          // the largest value first, then the smallest column among the lanes that hold it
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c0 + j;
            if (m >= rows) continue;  // warp-uniform
            uint32_t kv = 0u;  // ordered bits of the value; 0 = no candidate
            if (n_ok) {
              const uint32_t b = __float_as_uint(v[j]);
              kv = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
            }
            const uint32_t best_v = __reduce_max_sync(L3_FULL, kv);
            const uint32_t best_n = __reduce_min_sync(L3_FULL, (n_ok && kv == best_v) ? (uint32_t)n : 0xffffffffu);
            if (lane == 0 && best_n != 0xffffffffu)
              atomicMax(e.best + m, ((unsigned long long)best_v << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(e.col_offset + (int)best_n)));
          }
        }
      }
      if (!det && !live) {  // nothing was read: still hand the buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cf::TMEM_COLS));
  }
}

template <int KIND, int BNA, int EPI>
cudaError_t launch_sw_t(const TcGemmArgs& a, cudaStream_t s) {
  using Cf = SwCfg<KIND, BNA>;
  auto kern = gemm_swap_kernel<KIND, BNA, EPI>;
  static bool attr_done[16] = {false};
  static int n_sm[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&n_sm[dev & 15], cudaDevAttrMultiProcessorCount, dev);
    attr_done[dev & 15] = true;
  }
  const bool b16 = KIND == TC_BF16;
  const CUtensorMap* W0 = tc_get_map(a.W[0], b16, a.N, a.K, Cf::BMW);
  const CUtensorMap* X0 = tc_get_map(a.A[0], b16, a.rows, a.K, BNA);
  const CUtensorMap* W1 = Cf::PARTS == 2 ? tc_get_map(a.W[1], b16, a.N, a.K, Cf::BMW) : W0;
  const CUtensorMap* X1 = Cf::PARTS == 2 ? tc_get_map(a.A[1], b16, a.rows, a.K, BNA) : X0;
  if (!W0 || !X0 || !W1 || !X1) return cudaErrorInvalidValue;
  const int sms = n_sm[dev & 15] > 0 ? n_sm[dev & 15] : 148;
  const int tiles_w = (a.N + Cf::BMW - 1) / Cf::BMW;
  const int nkb = (a.K + Cf::BK - 1) / Cf::BK;
  // K-split (deterministic, see the epilogue) when the weight row blocks cover less than half the machine
  // (QKV: 48 blocks, Wo / Wdown: 32 blocks at the 8B shape): otherwise a few SMs stream the whole matrix at
  // their per-SM ingress rate.  Needs caller-owned scratch (TcGemmArgs::part / tile_cnt).
  int ksplit = 1;
  float* part = a.part;
  static const int resid_atomic = [] { const char* v = getenv("L3_SWAP_RESID_ATOMIC"); return v ? atoi(v) : 1; }();
  static const int det_epi = [] { const char* v = getenv("L3_SWAP_KSPLIT_DET"); return v ? atoi(v) : 1; }();
  if (EPI == EPI_RESID && b16 && a.e.out == a.e.resid && resid_atomic) {
    // bf16 mode residual epilogue: partials can be added in place with fp32 atomics (reorders the sum, which
    // the fp32 token-identical mode must not do) - no reduction tail at all
    while (tiles_w * ksplit * 2 <= sms && nkb / (ksplit * 2) >= 8) ksplit *= 2;
    part = nullptr;
  } else if (a.part && a.tile_cnt && tiles_w * 2 <= sms && tiles_w <= a.tile_cnt_len && (det_epi || EPI == EPI_RESID)) {
    static const int max_split = [] { const char* v = getenv("L3_SWAP_KSPLIT"); return v ? atoi(v) : 8; }();
    ksplit = std::min(std::min(max_split, sms / tiles_w), nkb / 8);
    const size_t need = (size_t)BNA * tiles_w * Cf::BMW * sizeof(float);
    while (ksplit > 1 && need * ksplit > a.part_bytes) --ksplit;
    while (ksplit > 1 && (ksplit - 1) * ((nkb + ksplit - 1) / ksplit) >= nkb) --ksplit;
    if (ksplit < 1) ksplit = 1;
  }
  // Measured and rejected at the 8B shape, batch 32 (round 2, profiles/r02_swap_gemm_experiments.txt): balanced K-ranges
  // (every CTA an equal contiguous range of (row block, k-block) units: gate|up 55.9 vs 54.2 us - the kernel is not
  // bound by the ragged second wave), a tile-contiguous copy of the weights read with one 16 KB bulk copy per stage
  // (5.06 vs 5.08 ms per step: not DRAM page locality either), a whole ring of weight stages requested ahead of the first
  // activation box (+2 us per GEMM: the first MMA then waits behind 160 KB of weights in the SM's TMA queue), and the
  // slices of a row block sharing its reduction + epilogue by rows behind a cooperative launch (QKV 37 vs 29 us).
  // (K-slices against wave quantisation when the row blocks exceed the SM count - gate|up at the 8B shape: 224 blocks
  // on 148 SMs - were measured and removed: 6.25 vs 5.57 ms per 8B batch-32 decode step, the reduction tail costs
  // more than the idle half wave.)
  dim3 grid(std::min(tiles_w * ksplit, sms));
  return launch_k(kern, grid, dim3(192), (size_t)Cf::SMEM, s, *W0, *W1, *X0, *X1, a.rows, a.N, a.K, ksplit, part, a.tile_cnt, a.e);
}

template <int KIND, int BNA>
cudaError_t launch_sw_e(const TcGemmArgs& a, cudaStream_t s) {
  switch (a.epi) {
    case EPI_STORE: return launch_sw_t<KIND, BNA, EPI_STORE>(a, s);
    case EPI_RESID: return launch_sw_t<KIND, BNA, EPI_RESID>(a, s);
    case EPI_SWIGLU: return launch_sw_t<KIND, BNA, EPI_SWIGLU>(a, s);
    case EPI_ARGMAX: return launch_sw_t<KIND, BNA, EPI_ARGMAX>(a, s);
    default: return launch_sw_t<KIND, BNA, EPI_ROPE_KV>(a, s);
  }
}
}  // namespace

bool gemm_swap_supported(int rows, int N) { return rows >= 1 && rows <= 128 && N >= 128; }

cudaError_t launch_gemm_swap(const TcGemmArgs& a, cudaStream_t s) {
  const int bna = a.rows <= 32 ? 32 : (a.rows <= 64 ? 64 : 128);
  if (a.kind == TC_BF16) {
    switch (bna) {
      case 32: return launch_sw_e<TC_BF16, 32>(a, s);
      case 64: return launch_sw_e<TC_BF16, 64>(a, s);
      default: return launch_sw_e<TC_BF16, 128>(a, s);
    }
  }
  switch (bna) {
    case 32: return launch_sw_e<TC_TF32X3, 32>(a, s);
    case 64: return launch_sw_e<TC_TF32X3, 64>(a, s);
    default: return launch_sw_e<TC_TF32X3, 128>(a, s);
  }
}
