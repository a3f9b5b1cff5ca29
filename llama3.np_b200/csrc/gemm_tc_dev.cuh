// Device side of the tcgen05 / TMA / TMEM GEMM (see gemm_tc.cu for the description): PTX wrappers, tile
// configuration and the warp-specialised pipeline as device functions, so that the stand-alone kernel
// (gemm_tc.cu) and the persistent batched-decode kernel (decode_batch.cu) run the very same code.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_tc.h"

#ifndef TC_STAMP
#define TC_STAMP(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------ PTX wrappers
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
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
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

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4)  // start address (16-byte units)
         | ((uint64_t)1 << 16)                   // leading byte offset (unused for swizzled K-major)
         | ((uint64_t)(1024 >> 4) << 32)         // stride byte offset: 8 rows x 128 B
         | ((uint64_t)1 << 46)                   // descriptor version (Blackwell)
         | ((uint64_t)2 << 61);                  // SWIZZLE_128B
}

template <int KIND, int BN> struct TcCfg {
  static constexpr int PARTS = KIND == TC_BF16 ? 1 : 2;
  static constexpr int BM = 128;
  static constexpr int BK = KIND == TC_BF16 ? 64 : 32;  // elements per 128-byte row
  static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = PARTS * (A_BYTES + B_BYTES);
  static constexpr int NACC = KIND == TC_TF32X3 ? 4 : (KIND == TC_TF32X3_2 ? 2 : 1);  // main accumulators + 1 correction, or 1
  static constexpr int NMAIN = NACC > 1 ? NACC - 1 : 1;
  static constexpr int ACC_COLS = NACC * BN;              // TMEM columns of one accumulator buffer
  static constexpr int NBUF = 2 * ACC_COLS <= 512 ? 2 : 1;  // double-buffered: epilogue(i) overlaps mainloop(i+1)
  static constexpr int TMEM_COLS = NBUF * ACC_COLS <= 32 ? 32 : NBUF * ACC_COLS <= 64 ? 64 : NBUF * ACC_COLS <= 128 ? 128
                                   : NBUF * ACC_COLS <= 256 ? 256 : 512;
  static constexpr int CH = BN < 64 ? BN : 64;            // epilogue column chunk
  static constexpr int LDC = CH + 2;
  static constexpr int EPI_BYTES = 4 * 32 * LDC * 4;      // per-warp staging tile [32 rows][CH + 2] fp32
  static constexpr int SMEM_MAX = 232448;
  static constexpr int STAGES_RAW = (SMEM_MAX - 1024 - 512 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 512 /*barriers*/;
  // instruction descriptor: D fp32, A/B bf16 or tf32, both K-major, N = BN, M = 128
  static constexpr uint32_t FMT = KIND == TC_BF16 ? 1u : 2u;
  static constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
};

// The pipeline state of one CTA: shared-memory ring + barriers, TMEM base, and - private to each
// role's thread - how many k-blocks (it) and tiles (ti) it has handled, so that a persistent caller can
// run several GEMMs back to back (tc_gemm_run) between one tc_pipe_setup / tc_pipe_teardown.
struct TcPipe {
  uint8_t* smem_raw;
  uint32_t raw, tiles, epi0, full0, empty0, accfull0, accempty0, tmem_slot, tmem_base;
  int nst;
  uint32_t it, ti;
};

// 192 threads: warp 0 = TMA, warp 1 = TMEM + MMA, warps 2-5 = epilogue.  smem: nst stages | EPI_BYTES | barriers.
template <int KIND, int BN>
__device__ __forceinline__ void tc_pipe_setup(TcPipe& p, uint8_t* smem_raw, int nst) {
  using Cf = TcCfg<KIND, BN>;
  p.smem_raw = smem_raw;
  p.raw = smem_u32(smem_raw);
  p.tiles = (p.raw + 1023u) & ~1023u;
  p.nst = nst;
  p.epi0 = p.tiles + nst * Cf::STAGE_BYTES;
  const uint32_t bars = p.epi0 + Cf::EPI_BYTES;  // full[nst] | empty[nst] | accfull[2] | accempty[2] | tmem ptr | flag
  p.full0 = bars; p.empty0 = bars + 8 * nst; p.accfull0 = bars + 16 * nst; p.accempty0 = p.accfull0 + 16;
  p.tmem_slot = p.accempty0 + 16;
  p.it = 0; p.ti = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(p.full0 + 8 * s, 1); mbar_init(p.empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(p.accfull0 + 8 * b, 1); mbar_init(p.accempty0 + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p.tmem_slot), "n"(Cf::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  p.tmem_base = *reinterpret_cast<uint32_t*>(smem_raw + (p.tmem_slot - p.raw));
}

template <int KIND, int BN>
__device__ __forceinline__ void tc_pipe_teardown(TcPipe& p) {
  using Cf = TcCfg<KIND, BN>;
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(p.tmem_base), "n"(Cf::TMEM_COLS));
  }
}


// One GEMM over the tiles first_tile, first_tile + tile_stride, ...; returns when this thread's role is done
// (the epilogue warps: all their global stores issued).
template <int KIND, int BN, int EPI>
__device__ __forceinline__ void tc_gemm_run(TcPipe& p, const CUtensorMap* tmA0p, const CUtensorMap* tmA1p,
                                            const CUtensorMap* tmB0p, const CUtensorMap* tmB1p, int rows, int N, int K,
                                            int a_box_rows, int ksplit, float* part, int* tile_cnt, const EpiArgs& e,
                                            int first_tile, int tile_stride) {
  using Cf = TcCfg<KIND, BN>;
  using KVT = typename std::conditional<KIND == TC_BF16, bf16, float>::type;
  constexpr int PARTS = Cf::PARTS, NBUF = Cf::NBUF, CH = Cf::CH, LDC = Cf::LDC;
  const int STAGES = p.nst;
  uint8_t* smem_raw = p.smem_raw;
  const uint32_t raw = p.raw, tiles = p.tiles, epi0 = p.epi0, full0 = p.full0, empty0 = p.empty0, accfull0 = p.accfull0,
                 accempty0 = p.accempty0, tmem_slot = p.tmem_slot, tmem_base = p.tmem_base;
  const CUtensorMap &tmA0 = *tmA0p, &tmA1 = *tmA1p, &tmB0 = *tmB0p, &tmB1 = *tmB1p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (rows + Cf::BM - 1) / Cf::BM, tiles_n = (N + BN - 1) / BN;
  const int ntiles_mn = tiles_m * tiles_n;
  const int ntiles = ntiles_mn * ksplit;  // K-split: tile t = (k slice t / ntiles_mn, output tile t % ntiles_mn)
  const int nkb = (K + Cf::BK - 1) / Cf::BK;
  const int kb_per = (nkb + ksplit - 1) / ksplit;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0));
    if (PARTS == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1));
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1));
    }
  }

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      uint32_t it = p.it;  // k-blocks issued so far (ring position)
      for (int t = first_tile; t < ntiles; t += tile_stride) {
        const int tmn = t % ntiles_mn, ks = t / ntiles_mn;
        const int m0 = (tmn % tiles_m) * Cf::BM, n0 = (tmn / tiles_m) * BN;
        const int kb0 = ks * kb_per, kb1 = min(nkb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t st = tiles + s * Cf::STAGE_BYTES;
          // the A box holds only the rows that exist (decode batches of 9..127 rows): rows beyond it
          // keep stale shared memory, which only feeds accumulator rows the epilogue never reads
          mbar_expect_tx(full0 + 8 * s, PARTS * (a_box_rows * 128 + Cf::B_BYTES));
          tma_load_2d(st, &tmA0, kb * Cf::BK, m0, full0 + 8 * s);
          tma_load_2d(st + PARTS * Cf::A_BYTES, &tmB0, kb * Cf::BK, n0, full0 + 8 * s);
          if (PARTS == 2) {
            tma_load_2d(st + Cf::A_BYTES, &tmA1, kb * Cf::BK, m0, full0 + 8 * s);
            tma_load_2d(st + PARTS * Cf::A_BYTES + Cf::B_BYTES, &tmB1, kb * Cf::BK, n0, full0 + 8 * s);
          }
          if (it < 12) TC_STAMP(2 + it);
        }
      }
      p.it = it;
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      uint32_t it = p.it, ti = p.ti;
      for (int t = first_tile; t < ntiles; t += tile_stride, ++ti) {
        const uint32_t buf = ti % NBUF, use = ti / NBUF;
        if (use > 0) {  // the epilogue must have drained this accumulator buffer
          mbar_wait(accempty0 + 8 * buf, (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + buf * Cf::ACC_COLS;
        const int ks = t / ntiles_mn;
        const int kb0 = ks * kb_per, kb1 = min(nkb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          if (it < 12) TC_STAMP(16 + it);
          const uint32_t st = tiles + s * Cf::STAGE_BYTES;
          const uint64_t a_hi = umma_desc_sw128(st), b_hi = umma_desc_sw128(st + PARTS * Cf::A_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 slices of 32 bytes along K inside the swizzle atom
            const uint64_t adv = (uint64_t)(kk * 2);
            const int slice = (kb - kb0) * 4 + kk;
            if (PARTS == 2) {
              const uint64_t a_lo = umma_desc_sw128(st + Cf::A_BYTES);
              const uint64_t b_lo = umma_desc_sw128(st + PARTS * Cf::A_BYTES + Cf::B_BYTES);
              const uint32_t corr = acc + Cf::NMAIN * BN, mainacc = acc + (slice % Cf::NMAIN) * BN;
              tc_mma<KIND>(corr, a_lo + adv, b_hi + adv, Cf::IDESC, slice == 0 ? 0u : 1u);
              tc_mma<KIND>(corr, a_hi + adv, b_lo + adv, Cf::IDESC, 1u);
              tc_mma<KIND>(mainacc, a_hi + adv, b_hi + adv, Cf::IDESC, slice < Cf::NMAIN ? 0u : 1u);
            } else {
              tc_mma<KIND>(acc, a_hi + adv, b_hi + adv, Cf::IDESC, slice == 0 ? 0u : 1u);
            }
          }
          tc_commit(empty0 + 8 * s);  // frees the stage once these MMAs have read it
        }
        tc_commit(accfull0 + 8 * buf);  // accumulator complete
        if (ti == 0) TC_STAMP(30);
      }
      p.it = it; p.ti = ti;
    }
  } else {  // ---------------- epilogue (warps 2-5)
    // Per 64-column chunk: TMEM -> registers (thread = accumulator row) -> this warp's private
    // shared-memory tile Cs[32][CH + 2]; then warp-wide rows: lanes own adjacent column pairs, so
    // every global access of the fused epilogue is coalesced.
    const int quarter = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) ..
    float* Cs = reinterpret_cast<float*>(smem_raw + (epi0 - raw)) + (warp - 2) * 32 * LDC;
    const int start_pos = (EPI == EPI_ROPE_KV) ? *e.pos_ptr : 0;
    uint32_t ti = p.ti;
    for (int t = first_tile; t < ntiles; t += tile_stride, ++ti) {
      const int tmn = t % ntiles_mn, ks = t / ntiles_mn;
      const int m0 = (tmn % tiles_m) * Cf::BM, n0 = (tmn / tiles_m) * BN;
      const uint32_t buf = ti % NBUF, use = ti / NBUF;
      mbar_wait(accfull0 + 8 * buf, use & 1);
      tc_fence_after();
      if (ti == 0 && threadIdx.x == 64) TC_STAMP(32);
      const uint32_t acc = tmem_base + buf * Cf::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
      const bool rows_live = m0 + quarter * 32 < rows;  // warp-uniform
      // accumulator columns c .. c+31 of this thread's row (3xTF32: the four accumulators summed)
      auto load_acc = [&](int c, float (&v)[32]) {
        tmem_ld32(acc + (uint32_t)c, v);
        if (Cf::NACC > 1) {  // the main accumulators plus the correction accumulator
#pragma unroll
          for (int k2 = 1; k2 < Cf::NACC; ++k2) {
            float w[32];
            tmem_ld32(acc + (uint32_t)c + k2 * BN, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += w[j];
          }
        }
      };
      const int Mp = tiles_m * Cf::BM, Np = tiles_n * BN;  // padded extents of the K-split scratch
      if (ksplit > 1) {
        // K-split: publish this slice's partial tile, then the LAST slice to arrive (per output tile)
        // sums all slices in slice order - a fixed order, so the result does not depend on timing -
        // and runs the fused epilogue on the sum.
        float* prow = part + ((size_t)ks * Mp + m0 + quarter * 32 + lane) * Np + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          if (n0 + c >= N || !rows_live) break;
          float v[32];
          load_acc(c, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) __stcg(reinterpret_cast<float4*>(prow + c + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
        // Variant for the next round (the pattern gemm_swap.cu already runs on hardware): the CTA barrier orders
        // the four warps' partial stores before thread 64's acq_rel atomic (cumulative release), whose acquire
        // side + the next barrier order the last arriver's reads - instead of two __threadfence() per thread
        // (MEMBAR + CCTL.IVALL: 12 % of this kernel's stall samples in the ncu source view).
        asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
        int* flag = reinterpret_cast<int*>(smem_raw + (tmem_slot - raw)) + 1;
        if (threadIdx.x == 64) {
          int old;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(tile_cnt + tmn) : "memory");
          *flag = (old == ksplit - 1);
          if (old == ksplit - 1) tile_cnt[tmn] = 0;  // ready for the next launch
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        const bool is_last = *flag != 0;
        asm volatile("bar.sync 2, 128;" ::: "memory");  // everyone has read the flag before it is reused
        if (!is_last) continue;
      }
      if constexpr (EPI == EPI_ARGMAX) {
        if (ksplit == 1) {
          // Greedy argmax fused into the LM head (llama3.py:320) without leaving the registers: a thread IS an
          // accumulator row, so it walks its row's columns in increasing order keeping (max, first index) and
          // merges into best[m] with ONE 64-bit atomicMax per tile.  (The shared-memory pass + per-row shuffle
          // tree this replaces took as long as the tile's 108-MMA main loop: VERDICT r01.)
          float bv = -INFINITY;
          int bi = 0x7fffffff;
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            if (n0 + c >= N || !rows_live) break;  // warp-uniform
            float v[32];
            load_acc(c, v);
            const int cbase = n0 + c;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j < N && v[j] > bv) { bv = v[j]; bi = cbase + j; }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
          const int m = m0 + quarter * 32 + lane;
          if (m < rows && bi != 0x7fffffff) atomicMax(e.best + m, argmax_key(bv, e.col_offset + bi));
          continue;
        }
      }
      int row_b = 0, row_pos = -1, row_real = 1;  // lane i: (sequence, position, not padding) of accumulator row quarter * 32 + i
      if (EPI == EPI_ROPE_KV && m0 + quarter * 32 + lane < rows) {
        const int m = m0 + quarter * 32 + lane;
        row_b = m / e.L;
        const int t = m - row_b * e.L;
        row_pos = (e.row_pos ? e.row_pos[row_b] : start_pos) + t;
        row_real = !e.row_len || t < e.row_len[row_b];
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        if (n0 + c0 >= N || !rows_live) break;  // warp-uniform
        __syncwarp();  // the previous chunk's reads of Cs are done
#pragma unroll
        for (int cc = 0; cc < CH; cc += 32) {
          float v[32];
          if (ksplit == 1) {
            load_acc(c0 + cc, v);
          } else {  // sum of the slices' partials, slice 0 first
            const float* src = part + ((size_t)(m0 + quarter * 32 + lane)) * Np + n0 + c0 + cc;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            // Variant for the next round (not yet measured; same additions in the same order, so bit-identical):
            // slice k2 + 1 is in flight while slice k2 is added - the loop below pays one L2 round trip per slice
            // (3-8 per tile at the stories15M shapes), this one about one per tile.
            float4 nxt[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) nxt[j] = __ldcg(reinterpret_cast<const float4*>(src + 4 * j));
            for (int k2 = 0; k2 < ksplit; ++k2) {
              float4 cur[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
              const size_t nk = (size_t)min(k2 + 1, ksplit - 1) * Mp * Np;  // clamped: the last prefetch is not used
#pragma unroll
              for (int j = 0; j < 8; ++j) nxt[j] = __ldcg(reinterpret_cast<const float4*>(src + nk + 4 * j));
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[4 * j] += cur[j].x; v[4 * j + 1] += cur[j].y; v[4 * j + 2] += cur[j].z; v[4 * j + 3] += cur[j].w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) *reinterpret_cast<float2*>(Cs + lane * LDC + cc + j) = make_float2(v[j], v[j + 1]);
        }
        if (ksplit == 1 && (c0 + CH >= BN || n0 + c0 + CH >= N)) {  // last TMEM read of this tile: hand the buffer back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
        }
        __syncwarp();
        const int cp = lane * 2;          // this lane's column pair inside the chunk
        const int col = n0 + c0 + cp;
        const bool col_ok = cp < CH && col < N;
        const bool has1 = col + 1 < N;
        if constexpr (EPI == EPI_RESID) {
          // x += tile: all 32 rows of this warp are loaded before the first store (one latency)
          if (col_ok) {
            float2 rr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int m = m0 + quarter * 32 + i;
              if (m < rows) rr[i] = *reinterpret_cast<const float2*>(e.resid + (size_t)m * e.ld_out + col);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int m = m0 + quarter * 32 + i;
              if (m < rows) {
                const float2 tt = *reinterpret_cast<const float2*>(Cs + i * LDC + cp);
                *reinterpret_cast<float2*>(e.out + (size_t)m * e.ld_out + col) = make_float2(rr[i].x + tt.x, rr[i].y + tt.y);
              }
            }
          }
        } else if constexpr (EPI == EPI_ARGMAX) {
          // greedy argmax fused into the LM head (llama3.py:320): per row, the chunk's best
          // (value, first index) is merged into best[m] with one 64-bit atomicMax.
#pragma unroll 1
          for (int i = 0; i < 32; ++i) {
            const int m = m0 + quarter * 32 + i;
            if (m >= rows) break;
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            if (col_ok) {
              const float2 tt = *reinterpret_cast<const float2*>(Cs + i * LDC + cp);
              bv = tt.x; bi = col;
              if (has1 && tt.y > bv) { bv = tt.y; bi = col + 1; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(L3_FULL, bv, o);
              const int oi = __shfl_xor_sync(L3_FULL, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0 && bi != 0x7fffffff) atomicMax(e.best + m, argmax_key(bv, e.col_offset + bi));
          }
        } else if constexpr (EPI == EPI_ROPE_KV) {
          // rotate q / k pairs (llama3.py:41-76) and append k, v to the cache (llama3.py:184-185).
          // (sequence, position) of the warp's 32 rows were computed once per tile (row_b, row_pos);
          // what depends on the column alone is computed once per chunk.
          const int qcols = e.HN * e.HD, kcols = e.KVHN * e.HD;
          const int region = col < qcols ? 0 : (col < qcols + kcols ? 1 : 2);
          const int within = region == 0 ? col : (region == 1 ? col - qcols : col - qcols - kcols);
          const int h = within / e.HD, d = within % e.HD, jj = d >> 1, hd2 = e.HD >> 1;
          KVT* cbase = (KVT*)(region == 1 ? e.cache_k : e.cache_v) + (size_t)h * e.M * e.HD + d;
          // The table entries of all 32 rows are requested before the first store: one L2 round trip per chunk
          // instead of one per group of four rows (a load may not move above a store that could alias it) - eight
          // dependent round trips per chunk made this epilogue (11 us per 128 x 256 tile) longer than a K = 2048
          // mainloop.
          float c[32], sn[32];
          const bool rot = region < 2 && col_ok;  // differs between lanes where a chunk straddles the q | k | v borders
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int pi = __shfl_sync(L3_FULL, row_pos, i);  // every lane takes part
            const size_t ti = (size_t)max(pi, 0) * hd2 + jj;   // rows beyond the batch: entry of position 0, unused
            c[i] = rot ? __ldg(e.cos_tab + ti) : 1.f;
            sn[i] = rot ? __ldg(e.sin_tab + ti) : 0.f;
          }
#pragma unroll
          for (int rb = 0; rb < 32; rb += 4) {
            if (m0 + quarter * 32 + rb >= rows) break;
            float2 v[4];
            int pos[4], bb[4], real[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pos[i] = __shfl_sync(L3_FULL, row_pos, rb + i);
              bb[i] = __shfl_sync(L3_FULL, row_b, rb + i);
              real[i] = __shfl_sync(L3_FULL, row_real, rb + i);
            }
            if (!col_ok) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const float2*>(Cs + (rb + i) * LDC + cp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (pos[i] < 0) continue;  // row beyond the batch
              float r0 = v[i].x, r1 = v[i].y;
              if (region < 2) {
                r0 = v[i].x * c[rb + i] - v[i].y * sn[rb + i];
                r1 = v[i].x * sn[rb + i] + v[i].y * c[rb + i];
              }
              if (region == 0) {
                const size_t o = (size_t)(m0 + quarter * 32 + rb + i) * e.ld_out + col;
                if (e.out) *reinterpret_cast<float2*>(e.out + o) = make_float2(r0, r1);
                if (e.out_bf16) *reinterpret_cast<__nv_bfloat162*>(e.out_bf16 + o) = __floats2bfloat162_rn(r0, r1);
              } else if (real[i]) {  // padding tokens of a ragged prefill leave the cache untouched
                KVT* ck = cbase + ((size_t)bb[i] * e.KVHN * e.M + pos[i]) * e.HD;
                if constexpr (sizeof(KVT) == 2) *reinterpret_cast<__nv_bfloat162*>(ck) = __floats2bfloat162_rn(r0, r1);
                else *reinterpret_cast<float2*>(ck) = make_float2(r0, r1);
              }
            }
          }
        } else if constexpr (EPI == EPI_SWIGLU) {
          // h = silu(gate) * up over interleaved (gate_j, up_j) columns (llama3.py:99-101).  bf16 mode
          // uses the fast exponential / reciprocal (its bar is 3e-2); fp32 mode keeps the exact form.
          if (col_ok) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
              const int m = m0 + quarter * 32 + i;
              if (m >= rows) break;
              const float2 t = *reinterpret_cast<const float2*>(Cs + i * LDC + cp);
              float hv;
              if constexpr (KIND == TC_BF16) hv = t.x * __fdividef(1.0f, 1.0f + __expf(-t.x)) * t.y;
              else hv = silu_ref(t.x) * t.y;
              const size_t o = (size_t)m * e.ld_out + (col >> 1);
              if (e.out_lo) { float hh, hl; split_tf32(hv, hh, hl); e.out[o] = hh; e.out_lo[o] = hl; }
              else if (e.out) e.out[o] = hv;
              if (e.out_bf16) e.out_bf16[o] = __float2bfloat16_rn(hv);
            }
          }
        } else {  // EPI_STORE
#pragma unroll 1
          for (int rb = 0; rb < 32; rb += 4) {
            if (m0 + quarter * 32 + rb >= rows) break;
            if (!col_ok) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + quarter * 32 + rb + i;
              const float2 t = *reinterpret_cast<const float2*>(Cs + (rb + i) * LDC + cp);
              if (m < rows) epilogue_pair<KVT>(EPI, e, m, col, t.x, t.y, has1);
            }
          }
        }
      }
      if (ksplit == 1 && (!rows_live || n0 >= N)) {  // nothing was read: still hand the buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
      }
      if (ti == 0 && threadIdx.x == 64) TC_STAMP(34);
    }
    p.ti = ti;
  }
}

