// Causal GQA prefill attention on the 5th-generation tensor cores (bf16 mode, head_dim 64 / 128):
// softmax(q k^T / sqrt(HD) + mask) v of llama3.py:190-207 for L > 1 as a flash-style kernel - the
// [L, T] score matrix (llama3.py:200-206) and the mask (llama3.py:293-297) are never materialised.
//
// One CTA = 128 queries of one (sequence, head), walking the visible keys in blocks of 128:
//   warp 0    TMA producer: Q tile once, then K and V blocks of the head's cache rows into
//             mbarrier-guarded shared-memory stages (128-byte swizzle)
//   warp 1    MMA issuer: S_j = Q K_j^T (tcgen05.mma, both operands K-major) into one of two TMEM
//             score buffers; O += P_j V_j with V as an MN-major operand straight from the cache
//             layout [key, head_dim] - no transposed copy.
//             Issue order S_0, S_1, PV_0, S_2, PV_1, ...: the tensor core computes the next
//             scores while the softmax warps work on the current block.
//   warps 2-5 online softmax, thread = query row (tcgen05.ld 32x32b: no shuffles): the 128 scores of a
//             row are read into registers with one wait, exp2 -> bf16 P is written to shared memory in
//             the K-major swizzled operand layout; the OUTPUT accumulates in TMEM across key blocks
//             (PV_j with accumulate) and is rescaled lazily - only when a row maximum has grown by
//             more than 2^8 (tcgen05.ld -> multiply -> tcgen05.st) - so a steady-state block costs the
//             softmax warps one TMEM wait instead of twelve.
// GQA (repeat_kv, llama3.py:79-83): q head h reads kv head h / n_rep.  The causal predicate
// key <= start_pos + t also hides cache rows beyond the prompt, so K/V boxes may overrun it.
#include <cuda.h>

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
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
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
// the same load without the wait (several can be in flight), and the wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major, 128-byte-swizzled operand (V: rows = keys = the MMA's K dimension, 128 bytes = 64
// head-dim elements per row): 8 key rows form a 1024-byte atom (stride byte offset between
// 8-key groups), 64-element head-dim atoms are lbo_bytes apart (leading byte offset).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int HD> struct FaCfg {
  static constexpr int BQ = 128, BKV = 128;
  static constexpr int QS = HD / 64;                      // 64-element (128-byte) slices of the head dim
  static constexpr int Q_BYTES = BQ * HD * 2;
  static constexpr int KV_BYTES = BKV * HD * 2;           // one K block or one V block
  static constexpr int P_BYTES = BQ * BKV * 2;
  static constexpr int STAGES = HD == 128 ? 2 : 4;
  static constexpr int TMEM_COLS = 512;                   // 2 x 128 score columns + 2 x HD output columns
  static constexpr int SMEM = Q_BYTES + 2 * STAGES * KV_BYTES + P_BYTES + 1024 + 256;
  // bf16 x bf16 -> fp32, M = 128; S: N = 128, both K-major;  PV: N = HD, B operand MN-major
  static constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((128u >> 4) << 24);
  static constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((128u >> 4) << 24);
};

template <int HD>
__global__ void __launch_bounds__(192, 1)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, AttnArgs a, int nrep) {
  using Cf = FaCfg<HD>;
  constexpr int STAGES = Cf::STAGES, QS = Cf::QS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sQ = (raw + 1023u) & ~1023u;
  const uint32_t sK = sQ + Cf::Q_BYTES, sV = sK + STAGES * Cf::KV_BYTES, sP = sV + STAGES * Cf::KV_BYTES;
  const uint32_t bars = sP + Cf::P_BYTES;
  // barrier map (8 bytes each)
  const uint32_t q_full = bars, k_full = bars + 8, k_empty = k_full + 8 * STAGES, v_full = k_empty + 8 * STAGES,
                 v_empty = v_full + 8 * STAGES, s_full = v_empty + 8 * STAGES, s_empty = s_full + 16,
                 pv_full = s_empty + 16, pv_empty = pv_full + 16, p_full = pv_empty + 16, p_empty = p_full + 8,
                 tmem_slot = p_empty + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  uint8_t* sP_ptr = smem_raw + (sP - raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;  // longest tiles (latest queries) first
  const int head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / nrep;
  const int q0 = qt * Cf::BQ;
  pdl_launch();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV));
    mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(k_empty + 8 * s, 1);
      mbar_init(v_full + 8 * s, 1); mbar_init(v_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full + 8 * i, 1); mbar_init(s_empty + 8 * i, 128);
      mbar_init(pv_full + 8 * i, 1); mbar_init(pv_empty + 8 * i, 128);
    }
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
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
  pdl_wait();
  const int start = *a.pos_ptr;
  const int last_q = min(q0 + Cf::BQ, a.L) - 1;             // last real query row of this tile
  const int nb = (start + last_q) / Cf::BKV + 1;            // key blocks holding a visible key
  const uint32_t tS = tmem_base, tO = tmem_base + 2 * Cf::BKV;  // two score buffers, then the output accumulator

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      mbar_expect_tx(q_full, Cf::Q_BYTES);
      for (int s = 0; s < QS; ++s) tma_load_2d(sQ + s * Cf::BQ * 128, &tmQ, head * HD + s * 64, b * a.L + q0, q_full);
      const int row0 = (b * a.KVHN + kvh) * a.M;
      for (int j = 0; j < nb; ++j) {
        const int st = j % STAGES, use = j / STAGES;
        if (use > 0) mbar_wait(k_empty + 8 * st, (use - 1) & 1);
        mbar_expect_tx(k_full + 8 * st, Cf::KV_BYTES);
        for (int s = 0; s < QS; ++s)
          tma_load_2d(sK + st * Cf::KV_BYTES + s * Cf::BKV * 128, &tmK, s * 64, row0 + j * Cf::BKV, k_full + 8 * st);
        if (use > 0) mbar_wait(v_empty + 8 * st, (use - 1) & 1);
        mbar_expect_tx(v_full + 8 * st, Cf::KV_BYTES);
        for (int s = 0; s < QS; ++s)
          tma_load_2d(sV + st * Cf::KV_BYTES + s * Cf::BKV * 128, &tmV, s * 64, row0 + j * Cf::BKV, v_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      mbar_wait(q_full, 0);
      tc_fence_after();
      auto issue_pv = [&](int j) {  // O += P_j V_j: the output accumulates in TMEM over all key blocks
        const int st = j % STAGES;
        mbar_wait(v_full + 8 * st, (j / STAGES) & 1);
        mbar_wait(p_full, j & 1);     // P_j written, and O rescaled if the softmax warps decided to
        tc_fence_after();
        const uint32_t vb = sV + st * Cf::KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < Cf::BKV / 16; ++kk) {  // 16 keys per MMA
          const uint64_t ad = desc_k_sw128(sP + (kk >> 2) * Cf::BQ * 128) + (uint64_t)((kk & 3) * 2);
          const uint64_t bd = desc_mn_sw128(vb + kk * 2048, Cf::BKV * 128);
          mma_bf16(tO, ad, bd, Cf::IDESC_PV, (j == 0 && kk == 0) ? 0u : 1u);
        }
        tc_commit(v_empty + 8 * st);
        tc_commit(p_empty);           // P_j consumed and O holds blocks 0..j
      };
      for (int j = 0; j < nb; ++j) {
        const int st = j % STAGES, buf = j & 1, use = j >> 1;
        mbar_wait(k_full + 8 * st, (j / STAGES) & 1);
        if (use > 0) mbar_wait(s_empty + 8 * buf, (use - 1) & 1);
        tc_fence_after();
        const uint32_t kb = sK + st * Cf::KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t adv = (uint64_t)((kk & 3) * 2);
          const uint64_t ad = desc_k_sw128(sQ + (kk >> 2) * Cf::BQ * 128) + adv;
          const uint64_t bd = desc_k_sw128(kb + (kk >> 2) * Cf::BKV * 128) + adv;
          mma_bf16(tS + buf * Cf::BKV, ad, bd, Cf::IDESC_S, kk == 0 ? 0u : 1u);
        }
        tc_commit(k_empty + 8 * st);
        tc_commit(s_full + 8 * buf);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(nb - 1);
    }
  } else {  // ---------------- softmax + output (warps 2-5): thread = query row
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                 // row inside the tile = TMEM lane
    const int qpos = start + q0 + r;                   // keys [0, qpos] are visible (llama3.py:293-297)
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float sc = 1.4426950408889634f / sqrtf((float)HD);  // log2(e) / sqrt(HD): p = 2^((s - m) * sc)
    // m_used: the row maximum the exponentials are taken against.  It follows the true running maximum
    // lazily: only when the maximum has grown by more than 2^8 is the TMEM output rescaled (and then for the
    // whole warp, tcgen05.ld / .st being warp-wide); until then p may reach 2^8, harmless in bf16 / fp32, and
    // O / l is exact either way because numerator and denominator share m_used.
    float m_used = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nb; ++j) {
      const int buf = j & 1, use = j >> 1;
      mbar_wait(s_full + 8 * buf, use & 1);
      tc_fence_after();
      const uint32_t ts = tS + buf * Cf::BKV + lane_off;
      const int key0 = j * Cf::BKV;
      const bool need_mask = key0 + Cf::BKV - 1 > qpos;
      // the whole score row in registers with ONE wait (the output no longer lives there)
      float v[Cf::BKV];
      tmem_ld32_issue(ts, v + 0); tmem_ld32_issue(ts + 32, v + 32);
      tmem_ld32_issue(ts + 64, v + 64); tmem_ld32_issue(ts + 96, v + 96);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty + 8 * buf);                       // S_j is in registers: the tensor core may overwrite it
      if (need_mask) {
#pragma unroll
        for (int i = 0; i < Cf::BKV; ++i)
          if (key0 + i > qpos) v[i] = -INFINITY;
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < Cf::BKV; i += 4) {
        mx4[0] = fmaxf(mx4[0], v[i]); mx4[1] = fmaxf(mx4[1], v[i + 1]);
        mx4[2] = fmaxf(mx4[2], v[i + 2]); mx4[3] = fmaxf(mx4[3], v[i + 3]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      bool waited = false;
      if (j == 0) {
        m_used = mx;                                          // block 0 always holds a visible key (key 0)
      } else {
        const bool grow = (mx - m_used) * sc > 8.0f;
        if (__any_sync(L3_FULL, grow)) {
          mbar_wait(p_empty, (j - 1) & 1);                    // PV_{j-1} done: O is stable
          waited = true;
          tc_fence_after();
          const float alpha = grow ? fast_exp2((m_used - mx) * sc) : 1.0f;
          if (grow) { m_used = mx; l_run *= alpha; }
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            float ov[32];
            tmem_ld32(tO + lane_off + c, ov);
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] *= alpha;
            tmem_st32(tO + lane_off + c, ov);
          }
          tc_fence_before();
        }
      }
      // exponentials first, into packed bf16 registers: this work overlaps PV_{j-1} on the tensor core; only
      // the shared-memory stores wait for it to release the P tile
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
      const float moff = m_used * sc;
      uint32_t pk[Cf::BKV / 2];
#pragma unroll
      for (int i = 0; i < Cf::BKV; i += 2) {
        const float p0 = fast_exp2(fmaf(v[i], sc, -moff)), p1 = fast_exp2(fmaf(v[i + 1], sc, -moff));  // 2^-inf = 0
        rs4[(i >> 1) & 3] += p0 + p1;
        __nv_bfloat162 t = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&t);
      }
      if (j > 0 && !waited) mbar_wait(p_empty, (j - 1) & 1);  // PV_{j-1} has consumed the previous P
#pragma unroll
      for (int c = 0; c < Cf::BKV; c += 32) {
        // columns c .. c+31 of row r: 4 chunks of 16 bytes, swizzled inside the 128-byte row
        uint8_t* prow = sP_ptr + (c >> 6) * (Cf::BQ * 128) + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j8 = ((c & 63) >> 3) + q;
          const int k4 = (c >> 1) + 4 * q;
          *reinterpret_cast<uint4*>(prow + ((j8 ^ (r & 7)) << 4)) = make_uint4(pk[k4], pk[k4 + 1], pk[k4 + 2], pk[k4 + 3]);
        }
      }
      l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P (generic proxy) -> tensor core (async proxy)
      mbar_arrive(p_full);
    }
    mbar_wait(p_empty, (nb - 1) & 1);                        // the last PV has landed in O
    tc_fence_after();
    const float inv = 1.0f / l_run;
    bf16* dst = a.out_bf16 + ((size_t)(b * a.L + q0 + r) * a.HN + head) * HD;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      float ov[32];
      tmem_ld32(tO + lane_off + c, ov);
      if (q0 + r < a.L) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 t = __floats2bfloat162_rn(ov[d + 2 * i] * inv, ov[d + 2 * i + 1] * inv);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(dst + c + d) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cf::TMEM_COLS));
  }
}

template <int HD>
cudaError_t launch_hd(const AttnArgs& a, const bf16* q16, cudaStream_t s) {
  using Cf = FaCfg<HD>;
  auto kern = attn_prefill_tc_kernel<HD>;
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM);
    if (e != cudaSuccess) return e;
    done[dev & 15] = true;
  }
  const int cache_rows = a.cache_rows;  // maxB * KVHN * M rows of HD elements
  const CUtensorMap* mq = tc_get_map(q16, true, a.B * a.L, a.HN * HD, Cf::BQ);
  const CUtensorMap* mk = tc_get_map(a.cache_k, true, cache_rows, HD, Cf::BKV);
  const CUtensorMap* mv = tc_get_map(a.cache_v, true, cache_rows, HD, Cf::BKV);
  if (!mq || !mk || !mv) return cudaErrorInvalidValue;
  dim3 grid((a.L + Cf::BQ - 1) / Cf::BQ, a.HN, a.B);
  return launch_k(kern, grid, dim3(192), (size_t)Cf::SMEM, s, *mq, *mk, *mv, a, a.HN / a.KVHN);
}
}  // namespace

bool attn_prefill_tc_supported(int HD) { return (HD == 64 || HD == 128) && tc_gemm_supported(HD); }

// q16: [B*L, HN*HD] bf16 rotated queries; K/V: the bf16 caches; out: a.out_bf16 [B*L, HN*HD]
cudaError_t launch_attn_prefill_tc(const AttnArgs& a, const bf16* q16, cudaStream_t s) {
  if (a.HD == 64) return launch_hd<64>(a, q16, s);
  if (a.HD == 128) return launch_hd<128>(a, q16, s);
  return cudaErrorInvalidValue;
}
