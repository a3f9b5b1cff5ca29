// Batched decode (many short sequences of a small model) as ONE cluster-resident kernel per step: all layers
// of Llama.__call__ with L == 1 (llama3.py:285-304: embedding, RMSNorm, QKV + RoPE + KV append, attention,
// output projection, SwiGLU FFN, final norm) run inside thread-block clusters that never synchronise with each
// other.  Only the LM head (gemm_tc.cu, fused argmax) stays a grid-wide kernel.
//
// Why: sequences are independent across the whole layer stack (only attention looks back, and only within a
// sequence), so a step needs no grid-wide barrier at all.  The kernel-per-projection graph spent 83 % of the
// stories15M B = 256 step in 41 launches that each moved ~1 us of data; here a cluster of C CTAs
// (C = n_heads) owns S sequences for the whole step, split Megatron-style INSIDE the cluster:
//   CTA j:  q/k/v rows of head j  ->  attention of head j (all S sequences)  ->  Wo restricted to head j's
//           columns (partial sums)  -> reduce-scatter + all-gather over DSMEM ->  gate/up rows of FFN slice j
//           -> Wdown restricted to slice j's columns (partial sums) -> reduce-scatter + all-gather.
// Two exchanges per layer, both through distributed shared memory; activations never leave the cluster.
//
// Data movement: every CTA streams ITS slice of the layer's weights (pre-packed k-major slabs, 648 KB per layer
// at the stories15M shape) and the K / V rows of its head from L2 / HBM with cp.async.bulk into a ring of 8 x 18 KB
// stages.  Four producer warps issue the copies (one thread sustains only ~49 GB/s of bulk copies, four reach
// ~220 GB/s per SM: scripts/ubench, profiles/r02_ubench.jsonl); nine compute warps consume.  Weights and cached
// K / V do not depend on this step's activations, so the producers run ahead across phase changes.
//
// Math: plain fp32 FFMA.  Thread tile = 4 output features x ALL S sequences (48 accumulators) over a k-group (1/4 or
// 1/8 of each slab's k rows): per k row one float4 of weights from the k-major slab (quarter warps read 128
// contiguous bytes) and S / 4 broadcast float4s of activations from a k-major [K][S + 8] buffer (80-byte rows keep
// the k-groups' reads in different banks) feed 48 FFMAs; k-groups are the high lane bits and are summed with shuffles.  (tcgen05 is the wrong tool
// for S = 12 rows: an M = 128 MMA costs 128 cycles whatever N is - profiles/r02_mma_cost.jsonl - and the legacy
// mma.sync path at 3xTF32 is only 1.4x the FFMA peak.)
#include <stdio.h>

#include "common.cuh"
#include "stack.h"

#ifndef SK_ST
#define SK_ST 8
#endif
#ifndef SK_LPP
#define SK_LPP 1
#endif
#ifndef SK_NBATCH
#define SK_NBATCH (SK_ST == 8 ? 1 : 4)
#endif

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded waits: a protocol bug must surface as a launch failure, never as a hung GPU.
constexpr uint32_t SK_SPIN_LIMIT = 1u << 26;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > SK_SPIN_LIMIT) __trap();
  } while (!ok);
}
// Producer flavour: a producer mostly waits for a stage to drain; a tight try_wait loop would take issue slots from
// the compute warps of its scheduler (ncu: a third of all executed instructions were polls), so it sleeps between probes.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) {
      __nanosleep(100);
      if (++spins > (SK_SPIN_LIMIT >> 4)) __trap();
    }
  } while (!ok);
}
// the cluster-scope flavour: pairs with a remote mbarrier.arrive.release.cluster
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > SK_SPIN_LIMIT) __trap();
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int D_, int HD_, int C_, int FD_, int S_>
struct SCfg {
  static constexpr int D = D_, HD = HD_, C = C_, FD = FD_, S = S_;
  static constexpr int QS = S / 4;       // float4 quads per activation row: a thread tile covers all S sequences
  // floats per k row of a GEMM operand buffer (k-major [K][XLD], sequences in slots 0 .. S-1).  80-byte rows put
  // the 16-byte chunks that the k-groups of a warp read in one instruction (rows k, k + 1, .. k + 7) into eight
  // different bank groups, so the broadcast loads are conflict-free.
  static constexpr int XLD = S % 8 == 0 ? S + 4 : S + 8;
  static constexpr int RLD = S;          // floats per row of the exchange buffers (receive slots, residual slice)
  static constexpr int NCW = 9;          // compute warps
  static constexpr int NCOMP = NCW * 32;
  static constexpr int NPW = 4;          // producer warps
  static constexpr int LPP = SK_LPP;     // issuing lanes per producer warp.  The ~3 bulk copies per us a thread can issue are a
                                         // per-THREAD limit and lanes of one warp scale like separate warps in isolation
                                         // (r02_ubench.jsonl), but here 2 lanes bought nothing (836 vs 844 k tok/s) and 4 lanes
                                         // with 16 stages lost 10 %: the divergent lanes' back-off sleeps stall each other
  static constexpr int NISS = NPW * LPP; // issuers; unit n belongs to issuer n % NISS
  static constexpr int NTHREADS = NCOMP + NPW * 32 + 32;  // + one warp whose lane 0 prefetches the next layer's K / V into L2
  // Ring: ST stages of 144 KB / ST.  Measured on the headline (profiles/r02_stack_sweep.jsonl): 8 x 18 KB is faster than
  // 16 x 9 KB although the latter lets an attention warp own two stages (its next unit loads under the current one's
  // math: units phase 10.5 -> 8.4 us per layer) - one thread issues only ~3 bulk copies per us whatever their size
  // (profiles/r02_ubench.jsonl), so halving the copies halves what the four producers can feed the GEMM phases.
  static constexpr int ST = SK_ST;       // ring stages
  static constexpr int STAGE = 147456 / ST;  // bytes per stage
  static constexpr int TCH = STAGE / (2 * HD * 4);  // cache rows per attention unit: [TCH K rows | TCH V rows] in one stage
  static constexpr int NAW = 8;          // warps that own attention units (ST == NAW: an owner reuses ITS stage; ST == 2 NAW:
                                         // it alternates between two)
  static constexpr int NCHMAX = (264 + TCH - 1) / TCH;  // attention units per sequence at the longest context: M <= NCHMAX * TCH
  static constexpr int PLD = HD + 4;     // floats per partial attention result: o[HD], m, l, pad
  static constexpr int ITER = 32 / ST;   // k rows of a slab per k-group
  static constexpr int NBATCH = SK_NBATCH;  // slabs a compute warp takes per wait / fence / release sequence
  // the four projections as seen by one CTA: output features F and reduction length K
  static constexpr int FA = 3 * HD, KA = D;        // q | k | v rows of this CTA's head        (llama3.py:166-168)
  static constexpr int FB = D, KB = HD;            // Wo restricted to this head's columns     (llama3.py:211)
  static constexpr int FC = 2 * FD / C, KC = D;    // interleaved gate / up rows of FFN slice   (llama3.py:99-100)
  static constexpr int FE = D, KE = FD / C;        // Wdown restricted to the slice's columns   (llama3.py:102)
  static constexpr int DS = D / C;                 // residual-stream columns a CTA owns in the reduce-scatter
  // k-groups of a projection: the largest power of two (<= 8) such that its F / 4 thread tiles x G fit the compute warps
  static constexpr int gof(int F) { return (F / 4) * 8 <= NCOMP ? 8 : ((F / 4) * 4 <= NCOMP ? 4 : ((F / 4) * 2 <= NCOMP ? 2 : 1)); }
  static constexpr int GA = gof(FA), GB = gof(FB), GC = gof(FC), GE = gof(FE);
  static constexpr int CTA_FLOATS = KA * FA + KB * FB + KC * FC + KE * FE;
  // shared memory (bytes)
  static constexpr int OFF_XT = ST * STAGE;
  static constexpr int OFF_XRES = OFF_XT + D * XLD * 4;
  static constexpr int OFF_Q = OFF_XRES + DS * RLD * 4;
  static constexpr int OFF_KN = OFF_Q + S * HD * 4;
  static constexpr int OFF_VN = OFF_KN + S * HD * 4;
  static constexpr int OFF_CTX = OFF_VN + S * HD * 4;
  static constexpr int OFF_RECV = OFF_CTX + HD * XLD * 4;
  static constexpr int OFF_PART = OFF_RECV + C * DS * RLD * 4;
  // h (written by the gate/up epilogue, read by Wdown) lives in the attention partials' space: the partials are dead
  // once the merge has run, and the next layer's attention starts only after every warp has left the Wdown GEMM
  static constexpr int OFF_H = OFF_PART;
  static constexpr int PART_BYTES = S * NCHMAX * PLD * 4 > KE * XLD * 4 ? S * NCHMAX * PLD * 4 : KE * XLD * 4;
  static constexpr int OFF_CS = OFF_PART + PART_BYTES;
  static constexpr int OFF_RED = OFF_CS + HD * 4;
  static constexpr int OFF_RINV = OFF_RED + NCW * 16 * 4;
  static constexpr int OFF_TOK = OFF_RINV + 64;
  static constexpr int OFF_PROG = OFF_TOK + 64;   // progress word of the compute warps (read by the L2 prefetcher)
  static constexpr int OFF_BAR = OFF_PROG + 16;
  static constexpr int SMEM = OFF_BAR + (2 * ST + 4) * 8;
  static_assert(S % 4 == 0 && S <= 16, "a thread tile holds 4 x S accumulators");
  static_assert(D % C == 0 && FD % C == 0 && DS % 4 == 0 && HD % 16 == 0 && KE % 4 == 0, "slices are float4-aligned");
  static_assert(ITER * GA * FA * 4 <= STAGE && ITER * GB * FB * 4 <= STAGE && ITER * GC * FC * 4 <= STAGE && ITER * GE * FE * 4 <= STAGE,
                "a slab of ITER k rows per k-group fits one stage");
  static_assert(KA % (ITER * GA) == 0 && KB % (ITER * GB) == 0 && KC % (ITER * GC) == 0 && KE % (ITER * GE) == 0, "whole slabs");
  static_assert(GB == GE && 32 % GB == 0, "both row-parallel projections push with the same lane pattern");
  static_assert((ST == NAW || ST == 2 * NAW) && NAW <= NCW && ST % NISS == 0 && TCH % 8 == 0, "attention owners / producers");
  static_assert(OFF_XT % 16 == 0 && OFF_XRES % 16 == 0 && OFF_Q % 16 == 0 && OFF_CTX % 16 == 0 && OFF_H % 16 == 0 &&
                OFF_RECV % 16 == 0 && OFF_PART % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
  static_assert(SMEM <= 232448, "shared memory");
};

#define SK_STAMP(a, idx)                                                                 \
  do {                                                                                   \
    if ((a).dbg && threadIdx.x == 0 && (idx) < 64) (a).dbg[(size_t)blockIdx.x * 128 + (idx)] = gtime(); \
  } while (0)

// sub-phase stamps of layer 2 only (slots 96 ..)
#define SK_STAMP2(a, l, idx)                                                                  \
  do {                                                                                        \
    if ((a).dbg && threadIdx.x == 0 && (l) == 2) (a).dbg[(size_t)blockIdx.x * 128 + 96 + (idx)] = gtime(); \
  } while (0)

template <class Cf> struct Ring {
  uint32_t base, full0, empty0;
  __device__ __forceinline__ uint32_t full(uint32_t slot) const { return full0 + 8 * slot; }
  __device__ __forceinline__ uint32_t empty(uint32_t slot) const { return empty0 + 8 * slot; }
};

template <class Cf> __device__ __forceinline__ void comp_sync() { asm volatile("bar.sync 1, %0;" ::"n"(Cf::NCOMP) : "memory"); }

// Exchanges over DSMEM are synchronised by DATA-ARRIVAL mbarriers: every warp that stores into a peer's shared
// memory arrives (release.cluster) on that peer's mbarrier after its lanes' stores; the receiver waits
// (acquire.cluster) for the fixed number of arrivals.
//   pbar: the partial sums of one row-parallel projection have landed in my receive buffer
//         (the warps of every source CTA that hold thread tiles of my columns: pwarps(rank) x C);
//   gbar: every owner's slice of the new residual stream has landed in my k-major buffer (GWARPS warps x C owners).
// Reuse of the buffers is safe without a further handshake: a peer pushes the NEXT partial sums only after all
// gathers of this exchange reached it, and my gather stores are issued after my reads of the receive buffer; a peer
// gathers into my residual buffer only after all my pushes reached it, and those follow my last read of that buffer.
// Two mbarriers of each kind alternate, so an arrival for exchange e + 1 can never be counted towards exchange e.
template <class Cf> struct XBars {
  uint32_t pbar0, gbar0;  // shared addresses of pbar[2], gbar[2]
  uint32_t np, ng;        // exchanges waited for so far
  // Arrivals are per WARP and destination, not per thread: a release.cluster arrive waits for the issuing thread's
  // earlier remote stores to be acknowledged (one DSMEM round trip), so a thread that alternates store / arrive
  // pays a round trip per destination (measured 2.4 us for the six destinations of the gather), and hundreds of
  // arrivals serialise on the receiver's mbarrier.  Every lane stores first, the warp synchronises (bar.warp.sync
  // orders the lanes' stores before the elected lanes' release, which is cumulative), then one lane per destination arrives.
  static constexpr int FPW = 32 / Cf::GB;                 // thread tiles per warp of the row-parallel projections
  static constexpr int TPO = Cf::DS / 4;                  // thread tiles per owner
  static constexpr int pwarps(int owner) {                // warps of one source CTA that hold tiles of `owner`
    int n = 0;
    for (int w = 0; w < Cf::NCW; ++w)
      if (w * FPW < (owner + 1) * TPO && (w + 1) * FPW > owner * TPO) ++n;
    return n;
  }
  static constexpr int GWARPS = (Cf::DS * Cf::QS + 31) / 32;  // warps that run reduce_and_gather
  static constexpr int GCOUNT = GWARPS * Cf::C;
  // Only warp 0 polls (acquire.cluster); the CTA barrier that follows hands the visibility on to the other warps,
  // which sleep in hardware meanwhile instead of spending issue slots on try_wait.
  __device__ __forceinline__ void wait_p() {
    if ((threadIdx.x >> 5) == 0) mbar_wait_cluster(pbar0 + 8 * (np & 1), (np >> 1) & 1);
    np += 1;
    comp_sync<Cf>();
  }
  __device__ __forceinline__ void wait_g() {
    if ((threadIdx.x >> 5) == 0) mbar_wait_cluster(gbar0 + 8 * (ng & 1), (ng >> 1) & 1);
    ng += 1;
    comp_sync<Cf>();
  }
};

// Where a compute thread sits in a projection with G k-groups: a warp holds 32 / G thread tiles (4 output features
// each, ALL S sequences), the k-group is the HIGH part of the lane index - so the eight lanes of a quarter warp read
// 128 contiguous bytes of a weight row (conflict-free LDS.128) and the k-groups are summed with shuffles.
template <int F, int G> struct Tile {
  static constexpr int FPW = 32 / G, NFG = F / 4;
  int kg, fg;
  bool active;
  __device__ __forceinline__ Tile() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    kg = lane / FPW;
    fg = warp * FPW + lane % FPW;
    active = fg < NFG;
  }
};

// One projection phase of this CTA: acc[4][S] (features 4 fg .. 4 fg + 3, every sequence) = sum over k.  The
// weights arrive as K / (ITER G) slabs of [ITER G k rows][F] floats in consecutive ring units; xt is the k-major
// activation buffer [K][XLD].  Returns with the k-groups summed: every lane of a tile holds the full sums.
template <class Cf, int F, int K, int G>
__device__ __forceinline__ void gemm_phase(const Ring<Cf>& rg, const uint8_t* ring, const float* xt, uint32_t& n,
                                           float (&acc)[4][Cf::S], long long* wait_cycles = nullptr) {
  constexpr int S = Cf::S, QS = Cf::QS, ITER = Cf::ITER, KSLAB = ITER * G, NSLAB = K / KSLAB;
  const Tile<F, G> tl;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int s = 0; s < S; ++s) acc[f][s] = 0.f;
  const float* w0 = reinterpret_cast<const float*>(ring) + tl.kg * F + 4 * tl.fg;
  auto math = [&](int slab, uint32_t slot) {
    const float* w = w0 + (size_t)slot * (Cf::STAGE / 4);
    const float* x = xt + (slab * KSLAB + tl.kg) * Cf::XLD;
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + i * G * F);
#pragma unroll
      for (int q = 0; q < QS; ++q) {
        const float4 x4 = *reinterpret_cast<const float4*>(x + i * G * Cf::XLD + 4 * q);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[0][4 * q + e] = fmaf(w4.x, xv[e], acc[0][4 * q + e]);
          acc[1][4 * q + e] = fmaf(w4.y, xv[e], acc[1][4 * q + e]);
          acc[2][4 * q + e] = fmaf(w4.z, xv[e], acc[2][4 * q + e]);
          acc[3][4 * q + e] = fmaf(w4.w, xv[e], acc[3][4 * q + e]);
        }
      }
    }
  };
  // Slabs are taken NBATCH at a time: one wait / fence / release sequence per NBATCH x ITER x 48 FFMAs of a thread,
  // and the loads of a later slab can be issued under an earlier slab's math.
  constexpr int NB = Cf::NBATCH;
  for (int slab = 0; slab < NSLAB; slab += NB) {
    const int cnt = NSLAB - slab < NB ? NSLAB - slab : NB;
    const long long tw0 = wait_cycles ? clock64() : 0;
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (j < cnt) mbar_wait(rg.full((n + j) % Cf::ST), ((n + j) / Cf::ST) & 1);
    const long long tc0 = wait_cycles ? clock64() : 0;
    if (wait_cycles) *wait_cycles += tc0 - tw0;
    if (tl.active) {
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (j < cnt) math(slab + j, (n + j) % Cf::ST);
    }
    if (wait_cycles) {  // debug: cycles from "stages ready" to "last FFMA issued" (the asm pins the clock read behind the math)
      int dep = 0;
#pragma unroll
      for (int f = 0; f < 4; ++f)
        asm volatile("" : "+r"(dep) : "f"(acc[f][0]), "f"(acc[f][1]), "f"(acc[f][2]), "f"(acc[f][3]), "f"(acc[f][4]), "f"(acc[f][5]),
                     "f"(acc[f][6]), "f"(acc[f][7]), "f"(acc[f][8]), "f"(acc[f][9]), "f"(acc[f][10]), "f"(acc[f][11]));
      wait_cycles[4] += clock64() - tc0 + dep;
    }
    // The stages were read through the generic proxy (LDS) and will be overwritten through the async proxy
    // (cp.async.bulk): every reader orders its own reads before the release with a proxy fence.  Without it the
    // refill occasionally overtook a slow warp's loads (measured: 30 % of 24-token runs at B = 256 deviated
    // bitwise; none with the fence - profiles/r02_stack_race.txt).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < cnt) mbar_arrive_n(rg.empty((n + lane) % Cf::ST), 1);  // this warp is done with the stages
    n += (uint32_t)cnt;
  }
#pragma unroll
  for (int off = Tile<F, G>::FPW; off < 32; off <<= 1)
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int s = 0; s < S; ++s) acc[f][s] += __shfl_xor_sync(L3_FULL, acc[f][s], off);
}

// RMSNorm (llama3.py:111-114) of the S residual rows, in place on the k-major buffer xt[D][XLD].
template <class Cf>
__device__ __forceinline__ void rms_inplace(float* xt, const float* __restrict__ g, float eps, float* red, float* rinv) {
  constexpr int S = Cf::S, D = Cf::D, QS = Cf::QS;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  static_assert(D <= 2 * Cf::NCOMP, "a thread owns at most two k rows");
  // the norm weights come from L2: ask for them before the sum of squares, not between the two barriers
  const float g0 = t < D ? g[t] : 0.f, g1 = t + Cf::NCOMP < D ? g[t + Cf::NCOMP] : 0.f;
  float ss[S];
#pragma unroll
  for (int s = 0; s < S; ++s) ss[s] = 0.f;
  for (int k = t; k < D; k += Cf::NCOMP) {
#pragma unroll
    for (int q = 0; q < QS; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(xt + k * Cf::XLD + 4 * q);
      ss[4 * q] = fmaf(v.x, v.x, ss[4 * q]); ss[4 * q + 1] = fmaf(v.y, v.y, ss[4 * q + 1]);
      ss[4 * q + 2] = fmaf(v.z, v.z, ss[4 * q + 2]); ss[4 * q + 3] = fmaf(v.w, v.w, ss[4 * q + 3]);
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s) ss[s] = warp_sum(ss[s]);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) red[warp * 16 + s] = ss[s];
  }
  comp_sync<Cf>();
  if (t < S) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < Cf::NCW; ++w) tot += red[w * 16 + t];
    rinv[t] = 1.0f / sqrtf(tot / (float)D + eps);
  }
  comp_sync<Cf>();
  for (int k = t; k < D; k += Cf::NCOMP) {
    const float gk = k == t ? g0 : g1;
#pragma unroll
    for (int q = 0; q < QS; ++q) {
      float4* p = reinterpret_cast<float4*>(xt + k * Cf::XLD + 4 * q);
      const float4 r4 = *reinterpret_cast<const float4*>(rinv + 4 * q);
      float4 v = *p;
      v.x = v.x * r4.x * gk; v.y = v.y * r4.y * gk; v.z = v.z * r4.z * gk; v.w = v.w * r4.w * gk;
      *p = v;
    }
  }
  comp_sync<Cf>();
}

// Partial sums of a row-parallel projection (Wo / Wdown restricted to this CTA's columns) go straight from the
// registers into slot [this rank] of the receive buffer of the CTA that owns the output columns.  The (feature,
// quad of sequences) items of a thread tile are dealt to its G lanes (every lane holds the full sums).
template <class Cf, int F, int G>
__device__ __forceinline__ void push_partials(const float (&acc)[4][Cf::S], uint32_t recv_local, uint32_t pbar_local, int rank) {
  const Tile<F, G> tl;
  static_assert(Tile<F, G>::NFG == Cf::NCW * Tile<F, G>::FPW && G == Cf::GB && F == Cf::D, "every lane of every compute warp holds a tile");
  const int f0 = 4 * tl.fg, owner = f0 / Cf::DS, fl = f0 % Cf::DS;  // DS % 4 == 0: the four features share an owner
  const uint32_t dst = mapa(recv_local, (uint32_t)owner) + (uint32_t)(((rank * Cf::DS + fl) * Cf::RLD) * 4);
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int q = 0; q < Cf::QS; ++q)
      if (((f + 4 * q) % G) == tl.kg)
        st_cluster_f4(dst + (uint32_t)((f * Cf::RLD + 4 * q) * 4), acc[f][4 * q], acc[f][4 * q + 1], acc[f][4 * q + 2], acc[f][4 * q + 3]);
  __syncwarp();
  // this warp's tiles belong to one owner or to two adjacent ones: lane 0 reports to the first, lane 1 to the second
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o_lo = (warp * Tile<F, G>::FPW * 4) / Cf::DS, o_hi = ((warp + 1) * Tile<F, G>::FPW * 4 - 1) / Cf::DS;
  static_assert(Tile<F, G>::FPW * 4 <= Cf::DS, "a warp's tiles span at most two owners");
  if (lane == 0) mbar_arrive_remote(mapa(pbar_local, (uint32_t)o_lo));
  if (lane == 1 && o_hi != o_lo) mbar_arrive_remote(mapa(pbar_local, (uint32_t)o_hi));
}

// The owner of residual columns [rank DS, rank DS + DS): x_new = x + sum over ranks (in rank order) of the
// partial sums (llama3.py:253 / :259), kept in xres and written into the k-major residual buffer of EVERY CTA.
template <class Cf>
__device__ __forceinline__ void reduce_and_gather(const float* recv, float* xres, uint32_t xt_local, uint32_t gbar_local, int rank) {
  constexpr int DS = Cf::DS;
  const int t = threadIdx.x, lane = t & 31;
  if ((t >> 5) >= XBars<Cf>::GWARPS) return;  // whole warps only: the arrival below follows a warp barrier
  if (t < DS * Cf::QS) {
    const int fl = t % DS, q = t / DS;
    float4* xr = reinterpret_cast<float4*>(xres + fl * Cf::RLD + 4 * q);
    float4 r = *xr;
#pragma unroll
    for (int p = 0; p < Cf::C; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(recv + (p * DS + fl) * Cf::RLD + 4 * q);
      r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
    }
    *xr = r;
    const uint32_t off = (uint32_t)(((rank * DS + fl) * Cf::XLD + 4 * q) * 4);
#pragma unroll
    for (int p = 0; p < Cf::C; ++p) st_cluster_f4(mapa(xt_local, (uint32_t)p) + off, r.x, r.y, r.z, r.w);
  }
  __syncwarp();
  if (lane < Cf::C) mbar_arrive_remote(mapa(gbar_local, (uint32_t)lane));
}

// debug dumps, [NL][4][B][D]: kind 0 = q, 1 = attention output, 2 / 3 = the residual stream after the first / second
// exchange of the layer (this CTA's columns)
template <class Cf>
__device__ __forceinline__ void dump_kmajor(const StackArgs& a, int l, int kind, int b0, int s_act, int col0, int ncols, int ld,
                                            const float* src) {
  for (int i = threadIdx.x; i < s_act * ncols; i += Cf::NCOMP) {
    const int s = i / ncols, c = i % ncols;
    a.dbg_x[(((size_t)l * 4 + kind) * a.B + b0 + s) * Cf::D + col0 + c] = src[c * ld + s];
  }
}
template <class Cf>
__device__ __forceinline__ void dump_rows(const StackArgs& a, int l, int kind, int b0, int s_act, int rank, const float* src_rowmajor) {
  for (int i = threadIdx.x; i < s_act * Cf::HD; i += Cf::NCOMP) {
    const int s = i / Cf::HD, d = i % Cf::HD;
    a.dbg_x[(((size_t)l * 4 + kind) * a.B + b0 + s) * Cf::D + rank * Cf::HD + d] = src_rowmajor[s * Cf::HD + d];
  }
}

template <class Cf>
__global__ void __launch_bounds__(Cf::NTHREADS, 1) decode_stack_kernel(const __grid_constant__ StackArgs a) {
  constexpr int D = Cf::D, HD = Cf::HD, C = Cf::C, S = Cf::S, QS = Cf::QS, XLD = Cf::XLD, RLD = Cf::RLD, TCH = Cf::TCH;
  // No integer round-trip on the base pointer: every shared-memory pointer below is plain pointer arithmetic on
  // the extern array, so the compiler keeps the address space and emits LDS / STS (32-bit addresses) instead of
  // generic LD / ST.  The kernel has no static shared memory: the dynamic window starts at offset 0.
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* base = smem_raw;
  uint8_t* ring = base;
  float* xt = reinterpret_cast<float*>(base + Cf::OFF_XT);      // [D][XLD]   residual stream, then norm(x), k-major
  float* xres = reinterpret_cast<float*>(base + Cf::OFF_XRES);  // [DS][RLD]  this CTA's columns of the residual stream
  float* q_s = reinterpret_cast<float*>(base + Cf::OFF_Q);      // [S][HD]    rotated q of this head
  float* kn_s = reinterpret_cast<float*>(base + Cf::OFF_KN);    // [S][HD]    this step's rotated k
  float* vn_s = reinterpret_cast<float*>(base + Cf::OFF_VN);    // [S][HD]    this step's v
  float* ctx_t = reinterpret_cast<float*>(base + Cf::OFF_CTX);  // [HD][XLD]  attention output of this head, k-major
  float* h_t = reinterpret_cast<float*>(base + Cf::OFF_H);      // [KE][XLD]  silu(gate) * up of this FFN slice, k-major
  float* recv = reinterpret_cast<float*>(base + Cf::OFF_RECV);  // [C][DS][RLD] partial sums from every rank
  float* part = reinterpret_cast<float*>(base + Cf::OFF_PART);  // [S][NCHMAX][PLD] attention partials
  float* cs = reinterpret_cast<float*>(base + Cf::OFF_CS);      // cos[HD/2] | sin[HD/2] of this position
  float* red = reinterpret_cast<float*>(base + Cf::OFF_RED);
  float* rinv = reinterpret_cast<float*>(base + Cf::OFF_RINV);
  int* tok = reinterpret_cast<int*>(base + Cf::OFF_TOK);
  volatile int* prog = reinterpret_cast<volatile int*>(base + Cf::OFF_PROG);  // layers whose attention has drained
  Ring<Cf> rg;
  rg.base = smem_u32(ring);
  rg.full0 = smem_u32(base + Cf::OFF_BAR);
  rg.empty0 = rg.full0 + 8 * Cf::ST;
  XBars<Cf> xb;
  xb.pbar0 = rg.empty0 + 8 * Cf::ST;
  xb.gbar0 = xb.pbar0 + 16;
  xb.np = 0; xb.ng = 0;

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int rank = (int)cluster_ctarank();
  const int b0 = (int)(blockIdx.x / C) * S;           // first sequence of this cluster
  const int s_act = min(S, a.B - b0);                  // sequences that exist
  const int step = a.scal[1] + 1;                      // llama3.py:316-318: decode step i runs at pos = L + i
  const int pos = a.scal[2] + step;
  const int nch = (pos + TCH - 1) / TCH;               // attention units per sequence: cache rows [0, pos)

  SK_STAMP(a, 63);
  if (t == 0) {
    for (int s = 0; s < Cf::ST; ++s) { mbar_init(rg.full(s), 1); mbar_init(rg.empty(s), Cf::NCW); }
    uint32_t pcount = 0;  // warps of every source CTA that hold tiles of my columns
#pragma unroll
    for (int o = 0; o < C; ++o)
      if (o == rank) pcount = (uint32_t)(XBars<Cf>::pwarps(o) * C);
    mbar_init(xb.pbar0, pcount); mbar_init(xb.pbar0 + 8, pcount);
    mbar_init(xb.gbar0, XBars<Cf>::GCOUNT); mbar_init(xb.gbar0 + 8, XBars<Cf>::GCOUNT);
    *prog = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist and its shared memory may be written from now on

  if (warp >= Cf::NCW) {
    // ================================================================== producers (LPP lanes of each producer warp)
    // All issuers walk the same unit sequence; issuer p takes the units n with n % NISS == p.  ST % NISS == 0, so
    // the previous use of a unit's stage was issued by the same thread: its wait on `empty` cannot alias.
    if (warp == Cf::NCW + Cf::NPW) {
      // ---------------------------------------------------------------- L2 prefetcher
      // The attention phases are HBM-bound in aggregate (every cluster streams its K / V at the same time) while
      // the projection phases between them leave HBM idle: as soon as layer l - 1's attention has drained the
      // ring, one thread asks L2 for the first pf_rows cache rows of layer l's K / V of this CTA's head
      // (cp.async.bulk.prefetch.L2: no destination, no completion), so the ring's copies find them in L2.
      if (lane == 0 && a.pf_rows > 0) {
        const uint32_t bytes = (uint32_t)min(pos, a.pf_rows) * HD * 4;
        for (int l = 1; l < a.NL; ++l) {
          uint32_t spins = 0;
          while (*prog < l) {
            __nanosleep(256);
            if (++spins > (1u << 24)) __trap();
          }
          const StackLayer ly = a.layers[l];
          for (int s = 0; s < s_act; ++s) {
            const size_t off = ((size_t)(b0 + s) * C + rank) * a.M * HD;
            for (uint32_t o = 0; o < bytes; o += 16384u) {
              const uint32_t nb = min(16384u, bytes - o);
              bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(ly.ck + off) + o, nb);
              bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(ly.cv + off) + o, nb);
            }
          }
        }
      }
    } else if (lane < Cf::LPP) {
      const int pw = (warp - Cf::NCW) + Cf::NPW * lane;  // issuer index
      uint32_t n = 0;
      const uint64_t pol_kv = policy_evict_first();
      bool kv_hint = false;
      auto issue = [&](const void* s0, uint32_t bytes0, const void* s1, uint32_t bytes1, uint32_t off1) {
        const uint32_t slot = n % Cf::ST, use = n / Cf::ST;
        if (use > 0) mbar_wait_backoff(rg.empty(slot), (use - 1) & 1);
        mbar_expect_tx(rg.full(slot), bytes0 + bytes1);
        const uint32_t dst = rg.base + slot * Cf::STAGE;
        if (kv_hint) {  // cached K / V is read once per step: first in line for eviction (weights and prefetched rows stay)
          bulk_g2s_hint(dst, s0, bytes0, rg.full(slot), pol_kv);
          bulk_g2s_hint(dst + off1, s1, bytes1, rg.full(slot), pol_kv);
        } else {
          bulk_g2s(dst, s0, bytes0, rg.full(slot));
          if (bytes1) bulk_g2s(dst + off1, s1, bytes1, rg.full(slot));
        }
      };
      auto slabs = [&](const float* w, int F, int K, int G) {
        const int ks = Cf::ITER * G;
        for (int sl = 0; sl < K / ks; ++sl, ++n)
          if ((int)(n % Cf::NISS) == pw) issue(w + (size_t)sl * ks * F, (uint32_t)(ks * F * 4), nullptr, 0u, 0u);
      };
      auto kv_units = [&](const StackLayer& ly) {
        uint32_t u = n;
        kv_hint = a.kv_evict_first != 0;
        for (int s = 0; s < s_act; ++s) {
          const size_t row0 = ((size_t)(b0 + s) * C + rank) * a.M;
          for (int c = 0; c < nch; ++c, ++u) {
            if ((int)(u % Cf::NISS) != pw) continue;
            const uint32_t bytes = (uint32_t)min(TCH, pos - c * TCH) * HD * 4;
            const size_t off = (row0 + (size_t)c * TCH) * HD;
            n = u;
            issue(ly.ck + off, bytes, ly.cv + off, bytes, (uint32_t)(TCH * HD * 4));
          }
        }
        n = u;
        kv_hint = false;
      };
      StackLayer ly = a.layers[0];
      for (int l = 0; l < a.NL; ++l) {
        StackLayer nx = ly;
        if (l + 1 < a.NL) nx = a.layers[l + 1];
        const float* w = ly.wpack + (size_t)rank * Cf::CTA_FLOATS;
        slabs(w, Cf::FA, Cf::KA, Cf::GA);
        w += Cf::KA * Cf::FA;
        kv_units(ly);
        slabs(w, Cf::FB, Cf::KB, Cf::GB);
        w += Cf::KB * Cf::FB;
        slabs(w, Cf::FC, Cf::KC, Cf::GC);
        w += Cf::KC * Cf::FC;
        slabs(w, Cf::FE, Cf::KE, Cf::GE);
        ly = nx;
      }
    }
    __syncwarp();
  } else {
    // ================================================================== compute warps
    uint32_t n = 0;
    const float scale = 1.0f / sqrtf((float)HD);
    // x = tok_embedding[token] (llama3.py:287), k-major; rows of sequences beyond the batch stay zero
    if (t < S) tok[t] = t < s_act ? a.d_next[b0 + t] : -1;
    if (t < HD / 2) {
      cs[t] = a.cos_tab[(size_t)pos * (HD / 2) + t];
      cs[HD / 2 + t] = a.sin_tab[(size_t)pos * (HD / 2) + t];
    }
    comp_sync<Cf>();
    for (int k = t; k < D; k += Cf::NCOMP) {
      float v[S];
#pragma unroll
      for (int s = 0; s < S; ++s) v[s] = tok[s] >= 0 ? a.embed[(size_t)tok[s] * D + k] : 0.f;
#pragma unroll
      for (int q = 0; q < QS; ++q)
        *reinterpret_cast<float4*>(xt + k * XLD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (k >= rank * Cf::DS && k < (rank + 1) * Cf::DS) {
#pragma unroll
        for (int q = 0; q < QS; ++q)
          *reinterpret_cast<float4*>(xres + (k - rank * Cf::DS) * RLD + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    comp_sync<Cf>();
    SK_STAMP(a, 0);

    float acc[4][S];
    for (int l = 0; l < a.NL; ++l) {
      const StackLayer ly = a.layers[l];
      long long wc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // debug: thread 0's cycles waiting for ring data [0..3] / in the math [4..7], per projection
      long long* wcp = a.dbg ? wc : nullptr;
      // ---- q, k, v = rope(norm(x) Wqkv^T) of head `rank`; k, v -> cache          llama3.py:248, 166-187
      rms_inplace<Cf>(xt, ly.norm_in, a.eps, red, rinv);
      SK_STAMP2(a, l, 0);
      gemm_phase<Cf, Cf::FA, Cf::KA, Cf::GA>(rg, ring, xt, n, acc, wcp);
      SK_STAMP2(a, l, 1);
      {
        // the tile's (feature pair, sequence) items are dealt to its G lanes: lane kg rotates pair kg & 1 of the
        // sequences s with s % (G / 2) == kg >> 1 (interleaved-pair rotation, llama3.py:41-76; v passes through)
        constexpr int G = Cf::GA;
        static_assert(G >= 2, "pair items need at least two lanes per tile");
        const Tile<Cf::FA, G> tl;
        if (tl.active) {
          const int pr = tl.kg & 1, sm = tl.kg >> 1;
          const int f0 = 4 * tl.fg, region = f0 / HD, d = f0 % HD + 2 * pr;  // region 0 q, 1 k, 2 v
          float c0 = 1.f, s0 = 0.f;
          if (region < 2) { c0 = cs[d >> 1]; s0 = cs[HD / 2 + (d >> 1)]; }
          float* dst_s = region == 0 ? q_s : (region == 1 ? kn_s : vn_s);
          float* cache = region == 1 ? ly.ck : ly.cv;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if ((s % (G / 2)) != sm) continue;
            const float v0 = pr ? acc[2][s] : acc[0][s], v1 = pr ? acc[3][s] : acc[1][s];
            const float2 o = make_float2(v0 * c0 - v1 * s0, v0 * s0 + v1 * c0);
            *reinterpret_cast<float2*>(dst_s + s * HD + d) = o;
            if (region > 0 && s < s_act)  // cache append at this position (llama3.py:184-185)
              *reinterpret_cast<float2*>(cache + (((size_t)(b0 + s) * C + rank) * a.M + pos) * HD + d) = o;
          }
        }
      }
      comp_sync<Cf>();
      if (a.dbg_x) dump_rows<Cf>(a, l, 0, b0, s_act, rank, q_s);
      if (l < 6) SK_STAMP(a, 1 + l * 10);
      // ---- ctx = softmax(q k^T / sqrt(HD)) v over keys [0, pos]                  llama3.py:190-207
      // Unit (s, c) = cache rows [c TCH, min(pos, (c + 1) TCH)) of sequence s, K rows then V rows in one stage,
      // owned by warp (s nch + c) % NAW: 4 lanes share a key, 8 keys per pass.  Each unit leaves a flash-style
      // partial (o, m, l); this step's own (k, v), still in shared memory, joins in the merge.
      {
        const uint32_t n0 = n;
        const int nunits = s_act * nch;
        n += (uint32_t)nunits;
        if (warp < Cf::NAW) {
          constexpr int CPL = HD / 16;  // 16-byte chunks per lane
          const int sub = lane >> 2, sl = lane & 3;
          for (int uu = warp; uu < nunits; uu += Cf::NAW) {
            const int s = uu / nch, c = uu - s * nch;
            const int rows = min(TCH, pos - c * TCH);
            const uint32_t nn = n0 + (uint32_t)uu, slot = nn % Cf::ST, use = nn / Cf::ST;
            float4 q4[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) q4[i] = *reinterpret_cast<const float4*>(q_s + s * HD + (sl + 4 * i) * 4);
            const long long ta0 = wcp ? clock64() : 0;
            mbar_wait(rg.full(slot), use & 1);
            const long long ta1 = wcp ? clock64() : 0;
            const float* Ks = reinterpret_cast<const float*>(ring + (size_t)slot * Cf::STAGE);
            const float* Vs = Ks + TCH * HD;
            constexpr int NP = TCH / 8;
            // Branch-free: rows beyond the unit's end are read from its last row (always valid data) and get weight
            // exp(-inf) = 0, so every load of a pass is independent of the predicate and the compiler can keep all
            // NP x CPL of them in flight (the predicated form serialised one LDS round trip per pass).
            float sc[NP];
            float mx = -INFINITY;
            {
              float4 k4[NP][CPL];
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const int rc = min(p * 8 + sub, rows - 1);
#pragma unroll
                for (int i = 0; i < CPL; ++i) k4[p][i] = *reinterpret_cast<const float4*>(Ks + rc * HD + (sl + 4 * i) * 4);
              }
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                float d = 0.f;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                  d = fmaf(q4[i].x, k4[p][i].x, d); d = fmaf(q4[i].y, k4[p][i].y, d);
                  d = fmaf(q4[i].z, k4[p][i].z, d); d = fmaf(q4[i].w, k4[p][i].w, d);
                }
                d += __shfl_xor_sync(L3_FULL, d, 1);
                d += __shfl_xor_sync(L3_FULL, d, 2);
                sc[p] = (p * 8 + sub) < rows ? d * scale : -INFINITY;
                mx = fmaxf(mx, sc[p]);
              }
            }
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 16));  // rows >= 1: finite
            float o[4 * CPL], lsum = 0.f;
#pragma unroll
            for (int e = 0; e < 4 * CPL; ++e) o[e] = 0.f;
            {
              float4 v4[NP][CPL];
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const int rc = min(p * 8 + sub, rows - 1);
#pragma unroll
                for (int i = 0; i < CPL; ++i) v4[p][i] = *reinterpret_cast<const float4*>(Vs + rc * HD + (sl + 4 * i) * 4);
              }
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const float pw_ = expf(sc[p] - mx);  // 0 for the rows beyond the end
                lsum += pw_;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                  o[4 * i] = fmaf(pw_, v4[p][i].x, o[4 * i]); o[4 * i + 1] = fmaf(pw_, v4[p][i].y, o[4 * i + 1]);
                  o[4 * i + 2] = fmaf(pw_, v4[p][i].z, o[4 * i + 2]); o[4 * i + 3] = fmaf(pw_, v4[p][i].w, o[4 * i + 3]);
                }
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async refill
            __syncwarp();
            if (lane == 0) mbar_arrive_n(rg.empty(slot), Cf::NCW);  // sole consumer of this stage
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
              lsum += __shfl_xor_sync(L3_FULL, lsum, off);
#pragma unroll
              for (int e = 0; e < 4 * CPL; ++e) o[e] += __shfl_xor_sync(L3_FULL, o[e], off);
            }
            if (wcp) { wc[8] += ta1 - ta0; wc[9] += clock64() - ta1; }

            if (sub == 0) {
              float* pp = part + (s * Cf::NCHMAX + c) * Cf::PLD;
#pragma unroll
              for (int i = 0; i < CPL; ++i)
                *reinterpret_cast<float4*>(pp + (sl + 4 * i) * 4) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
              if (sl == 0) { pp[HD] = mx; pp[HD + 1] = lsum; }
            }
          }
        }
        if (wcp && l == 2 && lane == 0 && warp < Cf::NAW) {  // slots 64 + 4 w ..: this warp's wait, math, finish time (last step)
          unsigned long long* dw = a.dbg + (size_t)blockIdx.x * 128 + 64 + 4 * warp;
          dw[0] = (unsigned long long)nunits; dw[1] = (unsigned long long)wc[8]; dw[2] = (unsigned long long)wc[9]; dw[3] = gtime();
        }
        comp_sync<Cf>();
        SK_STAMP2(a, l, 8);
        if (t == 0) *prog = l + 1;  // this layer's K / V has left the ring: the prefetcher may ask for the next layer's
        // merge: one thread per (sequence, 16-byte chunk of the head); every thread recomputes the score of this
        // step's own key (48 FMAs from broadcast reads) instead of sharing it through shuffles
        if (t < S * (HD / 4)) {
          constexpr int NCK = HD / 4;
          const int s = t / NCK, j = t % NCK;
          const int nc = s < s_act ? nch : 0;
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < NCK; ++i) {
            const float4 qa = *reinterpret_cast<const float4*>(q_s + s * HD + 4 * i);
            const float4 ka = *reinterpret_cast<const float4*>(kn_s + s * HD + 4 * i);
            d = fmaf(qa.x, ka.x, d); d = fmaf(qa.y, ka.y, d); d = fmaf(qa.z, ka.z, d); d = fmaf(qa.w, ka.w, d);
          }
          const float s_new = d * scale;
          const float* pp = part + s * Cf::NCHMAX * Cf::PLD;
          float pm[Cf::NCHMAX], pl[Cf::NCHMAX];
          float4 po[Cf::NCHMAX];
          float mx = s_new;
#pragma unroll
          for (int c = 0; c < Cf::NCHMAX; ++c) {
            const bool on = c < nc;
            pm[c] = on ? pp[c * Cf::PLD + HD] : -INFINITY;
            pl[c] = on ? pp[c * Cf::PLD + HD + 1] : 0.f;
            po[c] = on ? *reinterpret_cast<const float4*>(pp + c * Cf::PLD + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            mx = fmaxf(mx, pm[c]);
          }
          const float wn = expf(s_new - mx);
          float lsum = wn;
          const float4 vn = *reinterpret_cast<const float4*>(vn_s + s * HD + 4 * j);
          float4 o = make_float4(wn * vn.x, wn * vn.y, wn * vn.z, wn * vn.w);
#pragma unroll
          for (int c = 0; c < Cf::NCHMAX; ++c) {
            const float w = expf(pm[c] - mx);  // 0 for the units that do not exist
            lsum = fmaf(pl[c], w, lsum);
            o.x = fmaf(po[c].x, w, o.x); o.y = fmaf(po[c].y, w, o.y); o.z = fmaf(po[c].z, w, o.z); o.w = fmaf(po[c].w, w, o.w);
          }
          const float inv = 1.0f / lsum;
          ctx_t[(4 * j + 0) * XLD + s] = o.x * inv;
          ctx_t[(4 * j + 1) * XLD + s] = o.y * inv;
          ctx_t[(4 * j + 2) * XLD + s] = o.z * inv;
          ctx_t[(4 * j + 3) * XLD + s] = o.w * inv;
        }
        comp_sync<Cf>();
      }
      if (a.dbg_x) dump_kmajor<Cf>(a, l, 1, b0, s_act, rank * HD, HD, XLD, ctx_t);
      if (l < 6) SK_STAMP(a, 2 + l * 10);
      // ---- x += ctx Wo^T                                                        llama3.py:210-211, 253
      gemm_phase<Cf, Cf::FB, Cf::KB, Cf::GB>(rg, ring, ctx_t, n, acc, wcp ? wcp + 1 : nullptr);
      SK_STAMP2(a, l, 7);
      push_partials<Cf, Cf::FB, Cf::GB>(acc, smem_u32(recv), xb.pbar0 + 8 * (xb.np & 1), rank);
      if (l < 6) SK_STAMP(a, 3 + l * 10);
      xb.wait_p();  // all partial sums for my columns are here
      reduce_and_gather<Cf>(recv, xres, smem_u32(xt), xb.gbar0 + 8 * (xb.ng & 1), rank);
      xb.wait_g();  // the whole new residual stream is here
      if (a.dbg_x) dump_kmajor<Cf>(a, l, 2, b0, s_act, rank * Cf::DS, Cf::DS, RLD, xres);
      if (l < 6) SK_STAMP(a, 4 + l * 10);
      // ---- h = silu(norm(x) Wgate^T) * (norm(x) Wup^T) of FFN slice `rank`          llama3.py:256, 99-101
      rms_inplace<Cf>(xt, ly.norm_post, a.eps, red, rinv);
      SK_STAMP2(a, l, 2);
      gemm_phase<Cf, Cf::FC, Cf::KC, Cf::GC>(rg, ring, xt, n, acc, wcp ? wcp + 2 : nullptr);
      SK_STAMP2(a, l, 3);
      {
        // features 4 fg .. 4 fg + 3 = (gate, up) of h columns 2 fg and 2 fg + 1 of this slice; lane kg takes column
        // 2 fg + (kg & 1) of the sequences s with s % (G / 2) == kg >> 1
        constexpr int G = Cf::GC;
        static_assert(G >= 2, "pair items need at least two lanes per tile");
        const Tile<Cf::FC, G> tl;
        if (tl.active) {
          const int pr = tl.kg & 1, sm = tl.kg >> 1;
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if ((s % (G / 2)) != sm) continue;
            const float gt = pr ? acc[2][s] : acc[0][s], up = pr ? acc[3][s] : acc[1][s];
            h_t[(2 * tl.fg + pr) * XLD + s] = silu_ref(gt) * up;
          }
        }
      }
      comp_sync<Cf>();
      if (l < 6) SK_STAMP(a, 5 + l * 10);
      // ---- x += h Wdown^T                                                        llama3.py:102, 259
      gemm_phase<Cf, Cf::FE, Cf::KE, Cf::GE>(rg, ring, h_t, n, acc, wcp ? wcp + 3 : nullptr);
      SK_STAMP2(a, l, 4);
      push_partials<Cf, Cf::FE, Cf::GE>(acc, smem_u32(recv), xb.pbar0 + 8 * (xb.np & 1), rank);
      if (l < 6) SK_STAMP(a, 6 + l * 10);
      xb.wait_p();
      SK_STAMP2(a, l, 5);
      reduce_and_gather<Cf>(recv, xres, smem_u32(xt), xb.gbar0 + 8 * (xb.ng & 1), rank);
      SK_STAMP2(a, l, 6);
      xb.wait_g();
      if (l < 6) SK_STAMP(a, 7 + l * 10);
      if (a.dbg_x) dump_kmajor<Cf>(a, l, 3, b0, s_act, rank * Cf::DS, Cf::DS, RLD, xres);
      if (a.dbg && t == 0 && l < 5) {  // slots 8, 9, 10 of the layer's block: wait cycles of A + B, C, D
        a.dbg[(size_t)blockIdx.x * 128 + 8 + l * 10] = (unsigned long long)(wc[0] + wc[1]);
        a.dbg[(size_t)blockIdx.x * 128 + 9 + l * 10] = (unsigned long long)wc[2];
        a.dbg[(size_t)blockIdx.x * 128 + 10 + l * 10] = (unsigned long long)wc[3];
        if (l == 2) {  // layer 2 also reports its math cycles in the (unused) last block
          for (int i = 0; i < 4; ++i) a.dbg[(size_t)blockIdx.x * 128 + 58 + i] = (unsigned long long)wc[4 + i];
        }
      }
    }
    // ---- final norm (llama3.py:304); every CTA writes its DS columns of the LM head's operand rows
    rms_inplace<Cf>(xt, a.norm_final, a.eps, red, rinv);
    for (int i = t; i < s_act * (Cf::DS / 4); i += Cf::NCOMP) {
      const int s = i / (Cf::DS / 4), k4 = (i % (Cf::DS / 4)) * 4 + rank * Cf::DS;
      float4 hi, lo;
      split_tf32(xt[(k4 + 0) * XLD + s], hi.x, lo.x);
      split_tf32(xt[(k4 + 1) * XLD + s], hi.y, lo.y);
      split_tf32(xt[(k4 + 2) * XLD + s], hi.z, lo.z);
      split_tf32(xt[(k4 + 3) * XLD + s], hi.w, lo.w);
      *reinterpret_cast<float4*>(a.xlast_hi + (size_t)(b0 + s) * D + k4) = hi;
      *reinterpret_cast<float4*>(a.xlast_lo + (size_t)(b0 + s) * D + k4) = lo;
    }
    SK_STAMP(a, 62);
  }
  cluster_sync_all();  // nobody leaves while a peer may still write to (or arrive on) its shared memory
}

// ------------------------------------------------------------------------------------------ weight packing
// wpack[rank][phase][k][f] (k-major slabs, see SCfg): the producer copies contiguous ranges only.
template <class Cf>
__global__ void stack_pack_kernel(const float* __restrict__ wqkv, const float* __restrict__ wo, const float* __restrict__ w13,
                                  const float* __restrict__ w2, float* __restrict__ out) {
  constexpr int D = Cf::D, HD = Cf::HD, C = Cf::C, FD = Cf::FD;
  const size_t total = (size_t)C * Cf::CTA_FLOATS;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int rank = (int)(i / Cf::CTA_FLOATS);
    int r = (int)(i % Cf::CTA_FLOATS);
    float v;
    if (r < Cf::KA * Cf::FA) {
      const int k = r / Cf::FA, f = r % Cf::FA, region = f / HD, d = f % HD;  // wqkv rows: q heads | k heads | v heads
      v = wqkv[((size_t)(region * C + rank) * HD + d) * D + k];
    } else if ((r -= Cf::KA * Cf::FA) < Cf::KB * Cf::FB) {
      const int k = r / Cf::FB, f = r % Cf::FB;
      v = wo[(size_t)f * (C * HD) + rank * HD + k];
    } else if ((r -= Cf::KB * Cf::FB) < Cf::KC * Cf::FC) {
      const int k = r / Cf::FC, f = r % Cf::FC;
      v = w13[((size_t)rank * Cf::FC + f) * D + k];
    } else {
      r -= Cf::KC * Cf::FC;
      const int k = r / Cf::FE, f = r % Cf::FE;
      v = w2[(size_t)f * FD + rank * Cf::KE + k];
    }
    out[i] = v;
  }
}

__global__ void stack_finalize_kernel(unsigned long long* __restrict__ best, int B, int32_t* __restrict__ next_ids,
                                      int64_t* __restrict__ tokens, int stride, int* __restrict__ scal) {
  const int step = scal[1] + 1;
  for (int r = threadIdx.x; r < B; r += blockDim.x) {
    const unsigned long long k = best[r];
    best[r] = 0ull;
    const int idx = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
    next_ids[r] = idx;
    tokens[(size_t)r * stride + step] = (int64_t)idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    scal[1] = step;
    scal[0] = scal[2] + step;
  }
}

using Stories15M = SCfg<288, 48, 6, 768, 12>;

template <class Cf> cudaError_t prepare() {
  static bool done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev & 15]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(decode_stack_kernel<Cf>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM);
  if (e != cudaSuccess) return e;
  done[dev & 15] = true;
  return cudaSuccess;
}

template <class Cf> void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* at, int clusters, cudaStream_t s) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(clusters * Cf::C);
  cfg.blockDim = dim3(Cf::NTHREADS);
  cfg.dynamicSmemBytes = Cf::SMEM;
  cfg.stream = s;
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = Cf::C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
}

}  // namespace

bool decode_stack_supported(int D, int HN, int KVHN, int HD, int FD, int M) {
  using Cf = Stories15M;
  return D == Cf::D && HN == Cf::C && KVHN == Cf::C && HD == Cf::HD && FD == Cf::FD && M <= Cf::NCHMAX * Cf::TCH;
}

size_t decode_stack_pack_bytes(int, int, int, int) { return (size_t)Stories15M::C * Stories15M::CTA_FLOATS * sizeof(float); }

cudaError_t decode_stack_pack_layer(const float* wqkv, const float* wo, const float* w13, const float* w2, int, int, int, int,
                                    float* wpack, cudaStream_t s) {
  stack_pack_kernel<Stories15M><<<296, 256, 0, s>>>(wqkv, wo, w13, w2, wpack);
  return cudaGetLastError();
}

int decode_stack_seqs_per_cluster() { return Stories15M::S; }

int decode_stack_max_clusters() {
  using Cf = Stories15M;
  if (prepare<Cf>() != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  fill_cfg<Cf>(cfg, at, 1, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, decode_stack_kernel<Cf>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

cudaError_t launch_decode_stack(const StackArgs& a, int, int, int, int, cudaStream_t s) {
  using Cf = Stories15M;
  cudaError_t e = prepare<Cf>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute at[1];
  fill_cfg<Cf>(cfg, at, (a.B + Cf::S - 1) / Cf::S, s);
  return cudaLaunchKernelEx(&cfg, decode_stack_kernel<Cf>, a);
}

cudaError_t launch_stack_finalize(unsigned long long* best, int B, int32_t* next_ids, int64_t* tokens, int stride, int* scal,
                                  cudaStream_t s) {
  stack_finalize_kernel<<<1, 256, 0, s>>>(best, B, next_ids, tokens, stride, scal);
  return cudaGetLastError();
}
