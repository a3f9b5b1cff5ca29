// Batch-1 decode step as ONE persistent kernel (Llama.generate's per-token forward,
// llama3.py:285-321 with L == 1): the whole step - embedding, every layer's RMSNorm + QKV + RoPE
// + KV append, GQA attention, output projection, SwiGLU FFN, final norm, LM head and the greedy
// argmax - runs in a single launch of one CTA per SM.
//
// Why: batch-1 decode is HBM-bound (every weight byte is read once per token) but a kernel per
// projection leaves the HBM idle across every launch boundary, dependency wait and activation
// staging (measured: 8 us floor per GEMV launch; 1B-shaped decode reached 36 % of HBM peak).
// Here the weight stream never stops:
//   * warp 7 of every CTA is a TMA producer: one thread walks the CTA's slice of every weight
//     matrix of the whole step, in order, and copies it with cp.async.bulk into a ring of
//     MG_STAGES x 16 KB shared-memory stages guarded by full/empty mbarriers.  Weights do not
//     depend on activations, so the producer runs ahead through grid barriers and phase
//     changes; only ring capacity holds it back (160 KB per SM = 3.6 us of HBM time in flight).
//   * warps 0-6 consume: each stage belongs to one warp (round-robin), which reads it with
//     conflict-free 16-byte LDS, multiplies with the activation vector staged in shared memory
//     (RMSNorm fused into that staging) and finishes its row pairs with the fused epilogues
//     (RoPE + KV append / residual / SwiGLU / running argmax).
//   * phases are separated by a sense-reversing grid barrier (all CTAs are co-resident: one per
//     SM); activations between phases live in L2 and are read with ld.global.cg.
// Rows of a matrix are dealt to CTAs in contiguous blocks, so a CTA's slice is one contiguous
// byte range: short rows travel several pairs per stage in a single bulk copy, long rows
// (K * sizeof > 8 KB) are cut into chunks that the same warp consumes in order.
#include <stdio.h>

#include "attn_decode.cuh"
#include "common.cuh"
#include "mega.h"
#include "comm.h"

#ifndef MG_ATT_U
#define MG_ATT_U 2
#endif

namespace {

constexpr int MG_NW = 7;                        // consumer warps (+ 1 producer warp = 8 warps: 2 per
                                                // scheduler, so ptxas may use up to 255 registers)
constexpr int MG_CONS = MG_NW * 32;             // consumer threads
constexpr int MG_THREADS = MG_CONS + 32;        // + TMA producer warp
constexpr int MG_STAGE = 16 * 1024;
constexpr int MG_STAGES = 10;
constexpr int MG_XS_BYTES = MG_MAX_K * 4;
constexpr int MG_SMEM = MG_STAGES * MG_STAGE + MG_XS_BYTES + 2 * MG_STAGES * 8 + 256 + 128;

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
// Bounded waits: a protocol bug must surface as a launch failure, never as a hung GPU.
constexpr uint32_t MG_SPIN_LIMIT = 1u << 26;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > MG_SPIN_LIMIT) __trap();
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MG_CONS) : "memory"); }
struct ConsSync { __device__ __forceinline__ void operator()() const { cons_sync(); } };

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Debug timeline (l3_debug_mega_timeline): 512 slots per CTA; consumer slot = layer * 16 + event
// (layers < 24), producer slot = 384 + layer * 4 + matrix.  Written by one thread, only if enabled.
#define MG_STAMP(a, idx)                                                                   \
  do {                                                                                     \
    if ((a).dbg && (idx) < 512) (a).dbg[(size_t)blockIdx.x * 512 + (idx)] = gtime();       \
  } while (0)

// Grid barrier over the consumer warps of all CTAs (the producer warps never wait here): one
// monotonic arrival counter.  Barrier k of this launch completes when the counter reaches
// base + (k + 1) * gridDim.x, where `base` is the counter value the previous launch left behind
// (published in bar_gen by CTA 0 after its last barrier, so every CTA reads a stable value however
// late it starts).  Arrival is a single red.release (no returned value to wait for), the wait is an
// ld.acquire poll: about two L2 round trips after the last CTA arrives.  Wrap-safe compare.
struct GridBar {
  uint32_t target;  // counter value that completes the next barrier
};
__device__ __forceinline__ void grid_sync(const MegaArgs& a, GridBar& gb, int stamp) {
  cons_sync();
  if (threadIdx.x == 0) {
    MG_STAMP(a, stamp);       // every warp of this CTA has finished the phase
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar_cnt) : "memory");
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_gpu(a.bar_cnt) - gb.target) < 0)
      if (++spins > MG_SPIN_LIMIT) __trap();
    MG_STAMP(a, stamp + 1);   // barrier released
  }
  gb.target += gridDim.x;
  cons_sync();
}

// ---------------------------------------------------------------------------------- work plan
// How one weight matrix [N, K] is cut for this CTA; the producer and the consumers evaluate the
// same function, so they agree on the order of ring stages without communicating.
struct MgPlan {
  int p0, np;    // first row pair, number of row pairs of this CTA
  int PP;        // row pairs per stage (whole rows, K <= kcmax)
  int C, kcmax;  // chunks per row pair (K > kcmax), elements per row per stage
  int nunits;
};
__device__ __forceinline__ MgPlan mg_plan(int N, int K, int es) {
  MgPlan pl;
  const int npairs = N >> 1;
  pl.p0 = (int)((long long)blockIdx.x * npairs / gridDim.x);
  pl.np = (int)((long long)(blockIdx.x + 1) * npairs / gridDim.x) - pl.p0;
  pl.kcmax = MG_STAGE / (2 * es);
  if (K <= pl.kcmax) {
    pl.PP = min(32, pl.kcmax / K);
    pl.C = 1;
    pl.nunits = (pl.np + pl.PP - 1) / pl.PP;
  } else {
    pl.PP = 1;
    pl.C = (K + pl.kcmax - 1) / pl.kcmax;
    pl.nunits = pl.np * pl.C;
  }
  return pl;
}

struct Ring {
  uint32_t stages, full0, empty0;
};

// ---------------------------------------------------------------------------------- producer
// Walks this CTA's units of one matrix in ring order and hands each to `issue(n, src0, bytes0,
// src1, bytes1, off1)`: one contiguous range (several whole row pairs) or two (a chunk of each
// row of a pair; the second lands off1 bytes into the stage).
template <typename Issue>
__device__ __forceinline__ void walk_matrix(const void* W, int N, int K, int es, uint32_t& n, Issue&& issue) {
  const MgPlan pl = mg_plan(N, K, es);
  const char* Wb = (const char*)W;
  if (pl.C == 1) {
    for (int u = 0; u < pl.nunits; ++u, ++n) {
      const int pair0 = pl.p0 + u * pl.PP, cnt = min(pl.PP, pl.np - u * pl.PP);
      issue(n, Wb + (size_t)(2 * pair0) * K * es, (uint32_t)cnt * 2u * (uint32_t)K * (uint32_t)es, (const char*)nullptr, 0u, 0u);
    }
  } else {
    for (int j = 0; j * MG_NW < pl.np; ++j) {
      const int cnt = min(MG_NW, pl.np - j * MG_NW);
      for (int c = 0; c < pl.C; ++c) {
        const int k0 = c * pl.kcmax, kc = min(pl.kcmax, K - k0);
        for (int w = 0; w < cnt; ++w, ++n) {
          const int pair = pl.p0 + j * MG_NW + w;
          const uint32_t bytes = (uint32_t)kc * (uint32_t)es;
          issue(n, Wb + ((size_t)(2 * pair) * K + k0) * es, bytes, Wb + ((size_t)(2 * pair + 1) * K + k0) * es, bytes,
                (uint32_t)(pl.kcmax * es));
        }
      }
    }
  }
}

template <typename Issue>
__device__ __forceinline__ void walk_step(const MegaArgs& a, int es, bool stamp, Issue&& issue) {
  const int qkv_rows = (a.HN + 2 * a.KVHN) * a.HD;
  uint32_t n = 0;
  for (int l = 0; l < a.NL; ++l) {
    const MegaLayer& ly = a.layers[l];
    walk_matrix(ly.wqkv, qkv_rows, a.D, es, n, issue);
    if (stamp && l < 32) MG_STAMP(a, 384 + l * 4 + 0);
    walk_matrix(ly.wo, a.D, a.HN * a.HD, es, n, issue);
    if (stamp && l < 32) MG_STAMP(a, 384 + l * 4 + 1);
    walk_matrix(ly.w13, 2 * a.FD, a.D, es, n, issue);
    if (stamp && l < 32) MG_STAMP(a, 384 + l * 4 + 2);
    walk_matrix(ly.w2, a.D, a.FD, es, n, issue);
    if (stamp && l < 32) MG_STAMP(a, 384 + l * 4 + 3);
  }
  walk_matrix(a.lm_head, a.VS, a.D, es, n, issue);
}

// A consumer warp waits for ring unit n.  Stages of one slot are consumed by DIFFERENT warps, so a
// fast warp may reach unit n while the slot still holds (or still waits for) unit n - MG_STAGES:
// a bare parity wait on `full` would then alias the previous phase and return at once.  Waiting
// first for the previous occupant's release (`empty`, phase use - 1) pins the slot to this use;
// the in-order producer guarantees that no barrier can be two phases behind a waiter.
__device__ __forceinline__ void wait_full(const Ring& rg, uint32_t n) {
  const uint32_t slot = n % MG_STAGES, use = n / MG_STAGES;
  if (use > 0) mbar_wait(rg.empty0 + 8 * slot, (use - 1) & 1);
  mbar_wait(rg.full0 + 8 * slot, use & 1);
}

// ---------------------------------------------------------------------------------- consumer math
// dot products of two weight rows (shared memory, kc elements) with the staged activations
template <typename WT>
__device__ __forceinline__ void dot2(const WT* s0, const WT* s1, const float* xk, int kc, int lane, float& acc0,
                                     float& acc1) {
  constexpr int VEC = Vec16<WT>::N;
#pragma unroll 4
  for (int k = lane * VEC; k < kc; k += 32 * VEC) {
    const uint4 ra = *reinterpret_cast<const uint4*>(s0 + k);
    const uint4 rb = *reinterpret_cast<const uint4*>(s1 + k);
    float a0[VEC], a1[VEC];
    Vec16<WT>::unpack(ra, a0);
    Vec16<WT>::unpack(rb, a1);
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; v += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xk + k + v);
      p0 = fmaf(a0[v], xv.x, p0); p1 = fmaf(a1[v], xv.x, p1);
      p0 = fmaf(a0[v + 1], xv.y, p0); p1 = fmaf(a1[v + 1], xv.y, p1);
      p0 = fmaf(a0[v + 2], xv.z, p0); p1 = fmaf(a1[v + 2], xv.z, p1);
      p0 = fmaf(a0[v + 3], xv.w, p0); p1 = fmaf(a1[v + 3], xv.w, p1);
    }
    acc0 += p0;
    acc1 += p1;
  }
}

struct Best { float v; int i; };

// the fused epilogues for the single activation row of batch-1 decode (cf. epilogue_pair)
template <int EPI, typename KVT>
__device__ __forceinline__ void mg_epilogue(const MegaArgs& a, const MegaLayer& ly, int col, float v0, float v1, int pos,
                                            float2 resid, Best& best, unsigned tp_next) {
  if constexpr (EPI == EPI_RESID) {  // llama3.py:253, 259
    // the partial (rank 0 folds the residual in, `resid` is zero elsewhere) goes straight from the epilogue into
    // region [buf][rank] of EVERY rank (over NVLink; on one GPU the region is local), tagged with the exchange
    // number; the sum over ranks becomes the new x when the next phase stages its activations (stage_sum)
    {
      // two {value, epoch} words in one 16-byte store; each 8-byte half validates itself at the receiver
      const size_t off = ((size_t)(tp_next & 1) * a.tp_world + a.tp_rank) * a.ll_words + col;
      const uint4 w = make_uint4(__float_as_uint(resid.x + v0), tp_next, __float_as_uint(resid.y + v1), tp_next);
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < a.tp_world)
          asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a.peer_ll[p] + off), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
    }
  } else if constexpr (EPI == EPI_SWIGLU) {  // llama3.py:99-101, rows interleaved gate_j, up_j
    __stcg(a.h + (col >> 1), silu_ref(v0) * v1);
  } else if constexpr (EPI == EPI_ROPE_KV) {  // llama3.py:41-76, 184-185
    const int qcols = a.HN * a.HD, kcols = a.KVHN * a.HD;
    if (col < qcols + kcols) {
      const int within = col < qcols ? col : col - qcols;
      const int j = (within % a.HD) >> 1;
      const float c = a.cos_tab[(size_t)pos * (a.HD >> 1) + j], s = a.sin_tab[(size_t)pos * (a.HD >> 1) + j];
      const float r0 = v0 * c - v1 * s, r1 = v0 * s + v1 * c;
      if (col < qcols) {
        __stcg(reinterpret_cast<float2*>(a.q + col), make_float2(r0, r1));
      } else {
        const int h = within / a.HD, d = within % a.HD;
        KVT* ck = (KVT*)ly.ck + ((size_t)h * a.M + pos) * a.HD + d;
        ck[0] = from_f32<KVT>(r0);
        ck[1] = from_f32<KVT>(r1);
      }
    } else {
      const int within = col - qcols - kcols;
      const int h = within / a.HD, d = within % a.HD;
      KVT* cv = (KVT*)ly.cv + ((size_t)h * a.M + pos) * a.HD + d;
      cv[0] = from_f32<KVT>(v0);
      cv[1] = from_f32<KVT>(v1);
    }
  } else {  // EPI_ARGMAX: llama3.py:320, first maximum wins
    if (v0 > best.v || (v0 == best.v && col < best.i)) { best.v = v0; best.i = col; }
    if (v1 > best.v || (v1 == best.v && col + 1 < best.i)) { best.v = v1; best.i = col + 1; }
  }
}

// One projection phase for the consumer warps: y = W xs with the fused epilogue.
template <typename WT, typename KVT, int EPI>
__device__ __forceinline__ void consume_matrix(const MegaArgs& a, const MegaLayer& ly, const Ring& rg, const uint8_t* ring,
                                               const float* xs, int N, int K, int pos, uint32_t& nbase, Best& best,
                                               unsigned tp_next = 0) {
  const MgPlan pl = mg_plan(N, K, (int)sizeof(WT));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (pl.C == 1) {
    for (int u = warp; u < pl.nunits; u += MG_NW) {
      const uint32_t n = nbase + u, slot = n % MG_STAGES;
      const int pair0 = pl.p0 + u * pl.PP, cnt = min(pl.PP, pl.np - u * pl.PP);
      float2 resid = make_float2(0.f, 0.f);
      if (EPI == EPI_RESID && lane < cnt && a.tp_rank == 0)
        resid = __ldcg(reinterpret_cast<const float2*>(a.x + 2 * (pair0 + lane)));
      wait_full(rg, n);
      const WT* st = reinterpret_cast<const WT*>(ring + (size_t)slot * MG_STAGE);
      float my0 = 0.f, my1 = 0.f;
      for (int pr = 0; pr < cnt; ++pr) {
        float acc0 = 0.f, acc1 = 0.f;
        dot2<WT>(st + (size_t)(2 * pr) * K, st + (size_t)(2 * pr + 1) * K, xs, K, lane, acc0, acc1);
        acc0 = warp_sum(acc0);
        acc1 = warp_sum(acc1);
        if (lane == pr) { my0 = acc0; my1 = acc1; }
      }
      // generic-proxy reads (LDS) before the async-proxy refill (cp.async.bulk): each reader fences its own reads
      // (the same hazard was measured in decode_stack.cu: profiles/r02_stack_race.txt)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(rg.empty0 + 8 * slot);  // stage drained: the producer may refill it
      if (lane < cnt) mg_epilogue<EPI, KVT>(a, ly, 2 * (pair0 + lane), my0, my1, pos, resid, best, tp_next);
    }
  } else {
    for (int j = 0; j * MG_NW < pl.np; ++j) {
      const int cnt = min(MG_NW, pl.np - j * MG_NW);
      if (warp >= cnt) continue;
      const int pair = pl.p0 + j * MG_NW + warp;
      float2 resid = make_float2(0.f, 0.f);
      if (EPI == EPI_RESID && lane == 0 && a.tp_rank == 0) resid = __ldcg(reinterpret_cast<const float2*>(a.x + 2 * pair));
      float acc0 = 0.f, acc1 = 0.f;
      for (int c = 0; c < pl.C; ++c) {
        const uint32_t n = nbase + (uint32_t)(j * MG_NW * pl.C + c * cnt + warp), slot = n % MG_STAGES;
        const int k0 = c * pl.kcmax, kc = min(pl.kcmax, K - k0);
        wait_full(rg, n);
        const WT* st = reinterpret_cast<const WT*>(ring + (size_t)slot * MG_STAGE);
        dot2<WT>(st, st + pl.kcmax, xs + k0, kc, lane, acc0, acc1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(rg.empty0 + 8 * slot);
      }
      acc0 = warp_sum(acc0);
      acc1 = warp_sum(acc1);
      if (lane == 0) mg_epilogue<EPI, KVT>(a, ly, 2 * pair, acc0, acc1, pos, resid, best, tp_next);
    }
  }
  nbase += pl.nunits;
}

// Stage K fp32 activations into shared memory, optionally RMS-normalised (llama3.py:111-114).
// One bulk copy (L2 -> shared memory, a single round trip whatever K is) issued by thread 0 and
// awaited by all consumer warps on `xbar`; the source was written by other SMs earlier in this
// launch and is ordered by the grid barrier.  The embedding row of layer 0 (weight type, read-only)
// takes the register path.
struct XStage {
  uint32_t bar;     // mbarrier (shared address)
  uint32_t phase;   // uses so far
};
constexpr int MG_G_REGS = 4;  // norm-weight float4s a thread keeps in registers (covers D <= 4096)
struct NormW {
  float4 g[MG_G_REGS];
};
__device__ __forceinline__ void load_norm_w(NormW& nw, const float* norm_w, int K) {
#pragma unroll
  for (int i = 0; i < MG_G_REGS; ++i) {
    const int k = (threadIdx.x + i * MG_CONS) * 4;
    if (k < K) nw.g[i] = *reinterpret_cast<const float4*>(norm_w + k);
  }
}
__device__ __forceinline__ void rms_scale(float* xs, float* red, int K, const float* norm_w, const NormW& nw, float eps,
                                          float ss_thread) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float ss = warp_sum(ss_thread);
  if (lane == 0) red[warp] = ss;
  cons_sync();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < MG_NW; ++w) tot += red[w];
  const float rinv = 1.0f / sqrtf(tot / (float)K + eps);
#pragma unroll
  for (int i = 0; i < MG_G_REGS; ++i) {
    const int k = (tid + i * MG_CONS) * 4;
    if (k < K) {
      float4 v = *reinterpret_cast<const float4*>(xs + k);
      const float4 g = nw.g[i];
      v.x = v.x * rinv * g.x; v.y = v.y * rinv * g.y; v.z = v.z * rinv * g.z; v.w = v.w * rinv * g.w;
      *reinterpret_cast<float4*>(xs + k) = v;
    }
  }
  for (int k = (tid + MG_G_REGS * MG_CONS) * 4; k < K; k += MG_CONS * 4) {  // D > 4096: the tail reads g from L2
    float4 v = *reinterpret_cast<const float4*>(xs + k);
    const float4 g = *reinterpret_cast<const float4*>(norm_w + k);
    v.x = v.x * rinv * g.x; v.y = v.y * rinv * g.y; v.z = v.z * rinv * g.z; v.w = v.w * rinv * g.w;
    *reinterpret_cast<float4*>(xs + k) = v;
  }
}
__device__ __forceinline__ void stage_x(float* xs, float* red, XStage& st, const float* src, int K, const float* norm_w,
                                        float eps) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    // earlier generic-proxy accesses of xs (previous phase, attention scratch) before the async write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(st.bar, (uint32_t)K * 4u);
    bulk_g2s(smem_u32(xs), src, (uint32_t)K * 4u, st.bar);
  }
  NormW nw;
  if (norm_w) load_norm_w(nw, norm_w, K);  // overlaps the bulk copy's round trip
  mbar_wait(st.bar, st.phase & 1);
  st.phase += 1;
  if (norm_w) {
    float ss = 0.f;
    for (int k = tid * 4; k < K; k += MG_CONS * 4) {
      const float4 v = *reinterpret_cast<const float4*>(xs + k);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    rms_scale(xs, red, K, norm_w, nw, eps, ss);
  }
  cons_sync();
}
template <typename WT>
__device__ __forceinline__ void stage_embedding(float* xs, float* red, const WT* row, int K, const float* norm_w, float eps,
                                                float* publish) {
  const int tid = threadIdx.x;
  float ss = 0.f;
  for (int k = tid * 4; k < K; k += MG_CONS * 4) {
    float4 v;
    if constexpr (sizeof(WT) == 4) {
      v = *reinterpret_cast<const float4*>(row + k);
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(row + k);
      v = make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                      __uint_as_float(t.y & 0xffff0000u));
    }
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    *reinterpret_cast<float4*>(xs + k) = v;
    if (publish) __stcg(reinterpret_cast<float4*>(publish + k), v);
  }
  NormW nw;
  load_norm_w(nw, norm_w, K);
  rms_scale(xs, red, K, norm_w, nw, eps, ss);
  cons_sync();
}

// ---------------------------------------------------------------------------------- tensor parallel
// The sum over ranks after a row-parallel projection, inside the kernel, with the flag IN the data: every value
// travels as an 8-byte word {fp32 bits, exchange number} written by one store (8-byte stores are single-copy
// atomic, also over NVLink), into region [exchange & 1][sender] of every rank.  The receiver polls the words it
// needs until they carry the current exchange number and sums them in rank order (identical on every rank).
// Compared with "store, fence.sys, grid barrier, flag with release.sys, poll the flag" this is ONE NVLink one-way
// latency per exchange, and the grid barrier between the producing phase and the consuming phase disappears:
// a CTA starts the next phase as soon as the values it needs have arrived.
// Two buffers are enough: a rank can send exchange e + 2 only after it has received all of e + 1 from every
// rank, and a CTA sends its part of e + 1 only after it has finished reading e.
__device__ __forceinline__ uint4 ld_ll2(const unsigned long long* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// xs = sum over ranks of the received partials (= the new residual stream x), optionally RMS-normalised;
// CTA 0 also publishes x for rank 0's next residual epilogue
__device__ __forceinline__ void stage_sum(const MegaArgs& a, unsigned epoch, float* xs, float* red, int K, const float* norm_w,
                                          float eps) {
  const int tid = threadIdx.x;
  const unsigned long long* region = a.peer_ll[a.tp_rank] + (size_t)(epoch & 1) * a.tp_world * a.ll_words;
  float ss = 0.f;
  // No grid barrier precedes this staging any more, so nothing else separates the previous phase's reads of xs from
  // the writes below: a warp that has finished its row pairs must not overwrite the vector other warps of this CTA
  // are still multiplying with (measured without this barrier: rare wrong K / V rows, scripts/mega_cache_check.py).
  cons_sync();
  if (a.tp_world >= 4) {
    // Two levels: if every CTA summed all `world` regions itself, the chip would read world x K x 8 bytes x CTAs from
    // L2 per exchange (38 MB at TP 8, K = 4096: ~7 us, twice per layer).  Instead CTA c sums the float4 groups
    // g = c, c + grid, ... (rank order, identical everywhere), publishes them as tagged words of a LOCAL vector, and
    // every CTA polls that one vector.  A grid barrier lies between any two residual exchanges, so a fast CTA cannot
    // overwrite words a slow one is still waiting for; two buffers by parity on top.
    unsigned long long* xsum = a.ll_xsum + (size_t)(epoch & 1) * L3_LL_VEC;
    for (int g = blockIdx.x + tid * gridDim.x; g * 4 < K; g += MG_CONS * gridDim.x) {
      const int k = g * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r0 = 0; r0 < 8; r0 += 4) {
        if (r0 >= a.tp_world) break;
        uint4 lo[4], hi[4];
        uint32_t spins = 0;
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r0 + r < a.tp_world) {
              const unsigned long long* src = region + (size_t)(r0 + r) * a.ll_words + k;
              lo[r] = ld_ll2(src);
              hi[r] = ld_ll2(src + 2);
            }
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r0 + r < a.tp_world) ok = ok && lo[r].y == epoch && lo[r].w == epoch && hi[r].y == epoch && hi[r].w == epoch;
          if (!ok && ++spins > MG_SPIN_LIMIT) __trap();
        } while (!ok);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r0 + r < a.tp_world) {
            v.x += __uint_as_float(lo[r].x); v.y += __uint_as_float(lo[r].z);
            v.z += __uint_as_float(hi[r].x); v.w += __uint_as_float(hi[r].z);
          }
      }
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(xsum + k), "r"(__float_as_uint(v.x)), "r"(epoch),
                   "r"(__float_as_uint(v.y)), "r"(epoch) : "memory");
      asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(xsum + k + 2), "r"(__float_as_uint(v.z)), "r"(epoch),
                   "r"(__float_as_uint(v.w)), "r"(epoch) : "memory");
      __stcg(reinterpret_cast<float4*>(a.x + k), v);  // for rank 0's next residual epilogue
    }
    for (int k = tid * 4; k < K; k += MG_CONS * 4) {
      uint4 lo, hi;
      uint32_t spins = 0;
      do {
        lo = ld_ll2(xsum + k);
        hi = ld_ll2(xsum + k + 2);
        if (++spins > MG_SPIN_LIMIT) __trap();
      } while (lo.y != epoch || lo.w != epoch || hi.y != epoch || hi.w != epoch);
      const float4 v = make_float4(__uint_as_float(lo.x), __uint_as_float(lo.z), __uint_as_float(hi.x), __uint_as_float(hi.z));
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      *reinterpret_cast<float4*>(xs + k) = v;
    }
  } else
  for (int k = tid * 4; k < K; k += MG_CONS * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r0 = 0; r0 < 8; r0 += 4) {  // four ranks' words in flight together, summed in rank order
      if (r0 >= a.tp_world) break;
      uint4 lo[4], hi[4];
      uint32_t spins = 0;
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r0 + r < a.tp_world) {
            const unsigned long long* src = region + (size_t)(r0 + r) * a.ll_words + k;
            lo[r] = ld_ll2(src);
            hi[r] = ld_ll2(src + 2);
          }
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (r0 + r < a.tp_world) ok = ok && lo[r].y == epoch && lo[r].w == epoch && hi[r].y == epoch && hi[r].w == epoch;
        if (!ok && ++spins > MG_SPIN_LIMIT) __trap();  // a lost peer fails the launch instead of hanging the GPU
      } while (!ok);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (r0 + r < a.tp_world) {
          v.x += __uint_as_float(lo[r].x); v.y += __uint_as_float(lo[r].z);
          v.z += __uint_as_float(hi[r].x); v.w += __uint_as_float(hi[r].z);
        }
    }
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    *reinterpret_cast<float4*>(xs + k) = v;
    if (blockIdx.x == 0) __stcg(reinterpret_cast<float4*>(a.x + k), v);
  }
  if (norm_w) {
    NormW nw;
    load_norm_w(nw, norm_w, K);
    rms_scale(xs, red, K, norm_w, nw, eps, ss);
  }
  cons_sync();
}

// xs = a locally produced tagged vector (the attention output), no norm
__device__ __forceinline__ void stage_tagged(const unsigned long long* src, unsigned tag, float* xs, int K) {
  for (int k = threadIdx.x * 4; k < K; k += MG_CONS * 4) {
    uint4 lo, hi;
    uint32_t spins = 0;
    do {
      lo = ld_ll2(src + k);
      hi = ld_ll2(src + k + 2);
      if (++spins > MG_SPIN_LIMIT) __trap();
    } while (lo.y != tag || lo.w != tag || hi.y != tag || hi.w != tag);
    *reinterpret_cast<float4*>(xs + k) = make_float4(__uint_as_float(lo.x), __uint_as_float(lo.z), __uint_as_float(hi.x), __uint_as_float(hi.z));
  }
  cons_sync();
}

template <typename WT, int HD, int NREP>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  using KVT = WT;
  // plain pointer arithmetic on the extern array (no integer round-trip), so that the compiler keeps the shared
  // address space and the consumers' loads are LDS.128, not generic LD.E.128; no static shared memory in this
  // kernel, so the dynamic window starts at offset 0
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* ring = base;
  float* xs = reinterpret_cast<float*>(base + MG_STAGES * MG_STAGE);
  uint8_t* barmem = base + MG_STAGES * MG_STAGE + MG_XS_BYTES;
  float* red = reinterpret_cast<float*>(barmem + 2 * MG_STAGES * 8);
  Ring rg;
  rg.stages = smem_u32(ring);
  rg.full0 = smem_u32(barmem);
  rg.empty0 = rg.full0 + 8 * MG_STAGES;
  constexpr int ES = (int)sizeof(WT);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(red + 32), 1);  // activation staging barrier
    for (int s = 0; s < MG_STAGES; ++s) { mbar_init(rg.full0 + 8 * s, 1); mbar_init(rg.empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int qkv_rows = (a.HN + 2 * a.KVHN) * a.HD;
  if (threadIdx.x >= MG_CONS) {
    if (threadIdx.x == MG_CONS) {
      // ================================================================ TMA producer (warp 7)
      walk_step(a, ES, true, [&](uint32_t n, const char* s0, uint32_t b0, const char* s1, uint32_t b1, uint32_t off1) {
        const uint32_t slot = n % MG_STAGES, use = n / MG_STAGES;
        if (use > 0) mbar_wait(rg.empty0 + 8 * slot, (use - 1) & 1);
        mbar_expect_tx(rg.full0 + 8 * slot, b0 + b1);
        const uint32_t dst = rg.stages + slot * MG_STAGE;
        bulk_g2s(dst, s0, b0, rg.full0 + 8 * slot);
        if (s1) bulk_g2s(dst + off1, s1, b1, rg.full0 + 8 * slot);
      });
    }
    return;
  }

  // ================================================================== consumer warps
  const int tid = threadIdx.x;
  const int step = a.scal[1] + 1;          // llama3.py:316-318: decode step i runs at pos = L + i
  const int pos = a.scal[2] + step;
  const int token = a.d_next[0];
  uint32_t nbase = 0;
  Best best{-INFINITY, 0x7fffffff};
  XStage xst{smem_u32(red + 32), 0};
  GridBar gb{*reinterpret_cast<volatile unsigned*>(a.bar_gen) + gridDim.x};
  unsigned tp_epoch = *reinterpret_cast<volatile unsigned*>(a.epoch);  // exchanges so far (bar_gen: barrier counter value at launch)
  using ASm = AttnDecodeSmem<HD, NREP, MG_NW, KVT>;
  static_assert(sizeof(ASm) <= MG_XS_BYTES, "attention scratch aliases the activation buffer");
  ASm& asmem = *reinterpret_cast<ASm*>(xs);

  for (int l = 0; l < a.NL; ++l) {
    const MegaLayer& ly = a.layers[l];
    // ---- q, k, v = rope(norm(x) Wqkv^T); k, v -> cache                   llama3.py:248, 166-187
    if (l == 0)  // x = tok_embedding[token] (llama3.py:287); CTA 0 publishes the residual stream
      stage_embedding<WT>(xs, red, (const WT*)a.embed + (size_t)token * a.D, a.D, ly.norm_in, a.eps, blockIdx.x == 0 ? a.x : nullptr);
    else
      stage_sum(a, tp_epoch, xs, red, a.D, ly.norm_in, a.eps);
    if (tid == 0 && l < 24) MG_STAMP(a, l * 16 + 0);   // activations staged
    consume_matrix<WT, KVT, EPI_ROPE_KV>(a, ly, rg, ring, xs, qkv_rows, a.D, pos, nbase, best);
    grid_sync(a, gb, l < 24 ? l * 16 + 1 : 512);
    // ---- ctx = softmax(q k^T / sqrt(HD)) v over keys [0, pos]              llama3.py:190-207
    {
      AttnArgs at{};
      at.q = a.q; at.cache_k = ly.ck; at.cache_v = ly.cv;
      at.out_ll = a.ll_ctx; at.out_tag = tp_epoch + 1;  // tagged output: the projection below waits for the values, not for a barrier
      at.B = 1; at.L = 1; at.HN = a.HN; at.KVHN = a.KVHN; at.HD = HD; at.M = a.M;
      // decode attention is latency-bound per CTA (one DRAM round trip per pass over its keys), so a
      // head's keys are spread over as many CTAs as the grid offers, down to 8 keys per split
      // (measured: one CTA per kv head at T = 130 takes 12 us, 18 splits take 4 us)
      const int nsplit = max(1, min(a.nsplit, (pos + 1) / 8));
      at.part_o = a.part_o; at.part_ml = a.part_ml; at.nsplit = nsplit; at.counters = a.attn_cnt;
      const int ngrp = a.HN / NREP, nitems = ngrp * nsplit;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x)
        // MG_ATT_U key batches in flight per lane group; 4 instead of 2 measured no gain at a 2 k context (1B: 0.905 vs
        // 0.907 ms per step) and costs registers (8B: 3.29 vs 3.19 ms)
        attn_decode_item<HD, NREP, KVT, MG_NW, true, ConsSync, MG_ATT_U>(at, a.HN / a.KVHN, item % nsplit, item / nsplit, ngrp, 0, pos + 1,
                                                                  tid, asmem, ConsSync());
    }
    if (tid == 0 && l < 24) { MG_STAMP(a, l * 16 + 3); MG_STAMP(a, l * 16 + 4); }
    // ---- x += ctx Wo^T                                                    llama3.py:210-211, 253
    stage_tagged(a.ll_ctx, tp_epoch + 1, xs, a.HN * a.HD);
    if (tid == 0 && l < 24) MG_STAMP(a, l * 16 + 5);
    consume_matrix<WT, KVT, EPI_RESID>(a, ly, rg, ring, xs, a.D, a.HN * a.HD, pos, nbase, best, tp_epoch + 1);
    // ---- h = silu(norm(x) Wgate^T) * (norm(x) Wup^T)                      llama3.py:256, 99-101
    // no grid barrier: the staging waits for the tagged values themselves
    if (tid == 0 && l < 24) { MG_STAMP(a, l * 16 + 6); MG_STAMP(a, l * 16 + 7); }
    stage_sum(a, ++tp_epoch, xs, red, a.D, ly.norm_post, a.eps);
    if (tid == 0 && l < 24) MG_STAMP(a, l * 16 + 8);
    consume_matrix<WT, KVT, EPI_SWIGLU>(a, ly, rg, ring, xs, 2 * a.FD, a.D, pos, nbase, best);
    grid_sync(a, gb, l < 24 ? l * 16 + 9 : 512);
    // ---- x += h Wdown^T                                                   llama3.py:102, 259
    stage_x(xs, red, xst, a.h, a.FD, nullptr, 0.f);
    if (tid == 0 && l < 24) MG_STAMP(a, l * 16 + 11);
    consume_matrix<WT, KVT, EPI_RESID>(a, ly, rg, ring, xs, a.D, a.FD, pos, nbase, best, tp_epoch + 1);
    if (tid == 0 && l < 24) { MG_STAMP(a, l * 16 + 12); MG_STAMP(a, l * 16 + 13); }
    ++tp_epoch;  // summed by the next staging (layer l + 1 or the head)
  }
  // ---- next = argmax(norm(x) lm_head^T)                                   llama3.py:304-307, 320
  stage_sum(a, tp_epoch, xs, red, a.D, a.norm_final, a.eps);
  consume_matrix<WT, KVT, EPI_ARGMAX>(a, a.layers[0], rg, ring, xs, a.VS, a.D, pos, nbase, best);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(L3_FULL, best.v, o);
    const int oi = __shfl_xor_sync(L3_FULL, best.i, o);
    if (ov > best.v || (ov == best.v && oi < best.i)) { best.v = ov; best.i = oi; }
  }
  if ((tid & 31) == 0 && best.i != 0x7fffffff) atomicMax(a.d_best, argmax_key(best.v, a.tp_rank * a.VS + best.i));
  grid_sync(a, gb, 382);
  if (blockIdx.x == 0) {
    unsigned long long k = 0ull;
    if (tid == 0) {
      k = __ldcg(a.d_best);
      *a.d_best = 0ull;
    }
    if (a.tp_world > 1) {
      // vocabulary-sharded head: every rank sends its packed (value, global index) key to every rank as two
      // tagged words behind the vector of the same region, and takes the maximum
      const unsigned epoch = ++tp_epoch;
      const size_t off = ((size_t)(epoch & 1) * a.tp_world + a.tp_rank) * a.ll_words + L3_LL_VEC;
      if (tid == 0) {
        for (int p = 0; p < a.tp_world; ++p)
          asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(a.peer_ll[p] + off), "r"((unsigned)(k & 0xffffffffull)),
                       "r"(epoch), "r"((unsigned)(k >> 32)), "r"(epoch) : "memory");
        const unsigned long long* region = a.peer_ll[a.tp_rank] + (size_t)(epoch & 1) * a.tp_world * a.ll_words + L3_LL_VEC;
        k = 0ull;
        for (int r = 0; r < a.tp_world; ++r) {
          uint4 w;
          uint32_t spins = 0;
          do {
            w = ld_ll2(region + (size_t)r * a.ll_words);
            if (++spins > MG_SPIN_LIMIT) __trap();
          } while (w.y != epoch || w.w != epoch);
          const unsigned long long kr = ((unsigned long long)w.z << 32) | w.x;
          k = kr > k ? kr : k;
        }
      }
    }
    if (tid == 0) {
      *a.epoch = tp_epoch;
      const int idx = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
      a.d_next[0] = idx;
      a.d_tokens[step] = (int64_t)idx;
      a.scal[1] = step;
      a.scal[0] = pos;
      *a.bar_gen = gb.target - gridDim.x;  // = the counter now: base of the next launch
    }
  }
}

template <typename WT, int HD, int NREP>
cudaError_t launch_t(const MegaArgs& a, int grid, cudaStream_t s) {
  auto kern = decode_mega_kernel<WT, HD, NREP>;
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
    if (e != cudaSuccess) return e;
    done[dev & 15] = true;
  }
  // The software grid barrier needs every CTA resident at once: a cooperative launch makes the driver guarantee
  // that (another stream's kernel, an MPS client, or a second model on the device may be holding SMs), and an
  // oversubscribed grid fails the launch with cudaErrorCooperativeLaunchTooLarge instead of spinning into a trap.
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = MG_SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

template <typename WT>
cudaError_t launch_w(const MegaArgs& a, int nrep, int grid, cudaStream_t s) {
  const int key = a.HD * 16 + nrep;
  switch (key) {
    case 16 * 16 + 1: return launch_t<WT, 16, 1>(a, grid, s);
    case 16 * 16 + 4: return launch_t<WT, 16, 4>(a, grid, s);
    case 48 * 16 + 1: return launch_t<WT, 48, 1>(a, grid, s);
    case 48 * 16 + 2: return launch_t<WT, 48, 2>(a, grid, s);
    case 64 * 16 + 4: return launch_t<WT, 64, 4>(a, grid, s);
    case 128 * 16 + 4: return launch_t<WT, 128, 4>(a, grid, s);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

bool decode_mega_supported(int D, int HN, int KVHN, int HD, int FD, int VS) {
  const int nrep = HN / KVHN;
  const int key = HD * 16 + nrep;
  const bool combo = key == 16 * 16 + 1 || key == 16 * 16 + 4 || key == 48 * 16 + 1 || key == 48 * 16 + 2 ||
                     key == 64 * 16 + 4 || key == 128 * 16 + 4;
  const int maxk = std::max(std::max(D, HN * HD), FD);
  return combo && maxk <= MG_MAX_K && D % 8 == 0 && FD % 8 == 0 && (HN * HD) % 8 == 0 && VS % 2 == 0;
}

cudaError_t launch_decode_mega(const MegaArgs& a, bool w_bf16, int grid, cudaStream_t s) {
  const int nrep = a.HN / a.KVHN;
  return w_bf16 ? launch_w<bf16>(a, nrep, grid, s) : launch_w<float>(a, nrep, grid, s);
}
