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
// Math: plain fp32 FFMA.  Thread tile = 4 output features x TS sequences over a k-group (1/2 or 1/4 of each slab's
// k rows), weights read as float4 from the k-major slab (conflict-free), activations as broadcast float4/float2
// from a k-major [K][16] buffer; k-groups are lane bits and are summed with shuffles.  (tcgen05 is the wrong tool
// for S = 12 rows: an M = 128 MMA costs 128 cycles whatever N is - profiles/r02_mma_cost.jsonl - and the legacy
// mma.sync path at 3xTF32 is only 1.4x the FFMA peak.)
#include <stdio.h>

#include "common.cuh"
#include "stack.h"

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
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
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
  static constexpr int SG = 2;           // sequence groups: a thread tile covers TS = S / 2 sequences
  static constexpr int TS = S / SG;
  static constexpr int XLD = 16;         // floats per k row of a k-major activation buffer: group g at [8 g, 8 g + TS)
  static constexpr int NCW = 9;          // compute warps
  static constexpr int NCOMP = NCW * 32;
  static constexpr int NPW = 4;          // producer warps (one issuing thread each)
  static constexpr int NTHREADS = NCOMP + NPW * 32;
  static constexpr int ST = 8;           // ring stages
  static constexpr int STAGE = 18432;    // bytes per stage
  static constexpr int TCH = STAGE / (2 * HD * 4);  // cache rows per attention unit: [TCH K rows | TCH V rows] in one stage
  static constexpr int NAW = 8;          // warps that own attention units (<= ST: see the ring notes in the kernel)
  static constexpr int NCHMAX = 6;       // attention units per sequence at the longest context: M <= NCHMAX * TCH
  static constexpr int PLD = HD + 4;     // floats per partial attention result: o[HD], m, l, pad
  // the four projections as seen by one CTA: output features F and reduction length K
  static constexpr int FA = 3 * HD, KA = D;        // q | k | v rows of this CTA's head        (llama3.py:166-168)
  static constexpr int FB = D, KB = HD;            // Wo restricted to this head's columns     (llama3.py:211)
  static constexpr int FC = 2 * FD / C, KC = D;    // interleaved gate / up rows of FFN slice   (llama3.py:99-100)
  static constexpr int FE = D, KE = FD / C;        // Wdown restricted to the slice's columns   (llama3.py:102)
  static constexpr int DS = D / C;                 // residual-stream columns a CTA owns in the reduce-scatter
  static constexpr int gof(int F) { return NCOMP / ((F / 4) * SG) >= 4 ? 4 : (NCOMP / ((F / 4) * SG) >= 2 ? 2 : 1); }
  static constexpr int GA = gof(FA), GB = gof(FB), GC = gof(FC), GE = gof(FE);  // k-groups per phase
  static constexpr int CTA_FLOATS = KA * FA + KB * FB + KC * FC + KE * FE;
  // shared memory (bytes)
  static constexpr int OFF_XT = ST * STAGE;
  static constexpr int OFF_XRES = OFF_XT + D * XLD * 4;
  static constexpr int OFF_Q = OFF_XRES + DS * XLD * 4;
  static constexpr int OFF_KN = OFF_Q + S * HD * 4;
  static constexpr int OFF_VN = OFF_KN + S * HD * 4;
  static constexpr int OFF_CTX = OFF_VN + S * HD * 4;
  static constexpr int OFF_H = OFF_CTX + HD * XLD * 4;
  static constexpr int OFF_RECV = OFF_H + KE * XLD * 4;
  static constexpr int OFF_PART = OFF_RECV + C * DS * XLD * 4;
  static constexpr int OFF_CS = OFF_PART + S * NCHMAX * PLD * 4;
  static constexpr int OFF_RED = OFF_CS + HD * 4;
  static constexpr int OFF_RINV = OFF_RED + NCW * 16 * 4;
  static constexpr int OFF_TOK = OFF_RINV + 64;
  static constexpr int OFF_BAR = OFF_TOK + 64;
  static constexpr int SMEM = OFF_BAR + (2 * ST + 4) * 8 + 128 /* base alignment */;
  static_assert(S % SG == 0 && TS >= 4 && TS <= 8 && TS % 2 == 0, "thread tile covers 4, 6 or 8 sequences");
  static_assert(D % C == 0 && FD % C == 0 && DS % 4 == 0 && HD % 16 == 0 && KE % 4 == 0, "slices are float4-aligned");
  static_assert(8 * GA * FA * 4 <= STAGE && 8 * GB * FB * 4 <= STAGE && 8 * GC * FC * 4 <= STAGE && 8 * GE * FE * 4 <= STAGE,
                "a slab of 8 k rows per k-group fits one stage");
  static_assert(KA % (8 * GA) == 0 && KB % (8 * GB) == 0 && KC % (8 * GC) == 0 && KE % (8 * GE) == 0, "whole slabs");
  static_assert((FA / 4) * SG * GA <= NCOMP && (FB / 4) * SG * GB <= NCOMP && (FC / 4) * SG * GC <= NCOMP && (FE / 4) * SG * GE <= NCOMP, "threads");
  static_assert(NAW <= ST && NAW <= NCW, "attention owners");
  static_assert(SMEM <= 232448, "shared memory");
  __device__ static __forceinline__ int xslot(int s) { return (s / TS) * 8 + (s % TS); }
};

#define SK_STAMP(a, idx)                                                                 \
  do {                                                                                   \
    if ((a).dbg && threadIdx.x == 0 && (idx) < 64) (a).dbg[(size_t)blockIdx.x * 64 + (idx)] = gtime(); \
  } while (0)

template <class Cf> struct Ring {
  uint32_t base, full0, empty0;
  __device__ __forceinline__ uint32_t full(uint32_t slot) const { return full0 + 8 * slot; }
  __device__ __forceinline__ uint32_t empty(uint32_t slot) const { return empty0 + 8 * slot; }
};

template <class Cf> __device__ __forceinline__ void comp_sync() { asm volatile("bar.sync 1, %0;" ::"n"(Cf::NCOMP) : "memory"); }

// Exchanges over DSMEM are synchronised by DATA-ARRIVAL mbarriers: every thread that stores into a peer's shared
// memory arrives (release.cluster) on that peer's mbarrier right after its own stores, so the release covers
// exactly the stores it orders; the receiver's threads wait (acquire.cluster) for the fixed number of arrivals.
//   pbar: the partial sums of one row-parallel projection have landed in my receive buffer
//         ((DS / 4) SG pushing threads per source CTA, C sources);
//   gbar: every owner's slice of the new residual stream has landed in my k-major buffer (DS SG threads x C owners).
// Reuse of the buffers is safe without a further handshake: a peer pushes the NEXT partial sums only after all
// gathers of this exchange reached it, and my gather stores are issued after my reads of the receive buffer; a peer
// gathers into my residual buffer only after all my pushes reached it, and those follow my last read of that buffer.
// Two mbarriers of each kind alternate, so an arrival for exchange e + 1 can never be counted towards exchange e.
template <class Cf> struct XBars {
  uint32_t pbar0, gbar0;  // shared addresses of pbar[2], gbar[2]
  uint32_t np, ng;        // exchanges waited for so far
  static constexpr int PCOUNT = (Cf::DS / 4) * Cf::SG * Cf::C;
  static constexpr int GCOUNT = Cf::DS * Cf::SG * Cf::C;
};

// One projection phase of this CTA: acc[4][TS] (thread tile: features 4 fg .. 4 fg + 3, sequences of group sg)
// += sum over the k rows of this thread's k-group.  The weights arrive as K / (8 G) slabs of [8 G k rows][F]
// floats in consecutive ring units; xt is the k-major activation buffer [K][XLD].  Returns with the k-groups
// summed (valid in the lanes with kg == 0; every lane holds the same sum).
template <class Cf, int F, int K, int G>
__device__ __forceinline__ void gemm_phase(const Ring<Cf>& rg, const uint8_t* ring, const float* xt, uint32_t& n,
                                           float (&acc)[4][Cf::TS]) {
  constexpr int TS = Cf::TS, NFG = F / 4, KSLAB = 8 * G, NSLAB = K / KSLAB, NTHR = NFG * Cf::SG * G;
  const int t = threadIdx.x, lane = t & 31;
  const bool active = t < NTHR;
  const int kg = t % G, u = t / G, fg = u % NFG, sg = (u / NFG) % Cf::SG;
#pragma unroll
  for (int f = 0; f < 4; ++f)
#pragma unroll
    for (int s = 0; s < TS; ++s) acc[f][s] = 0.f;
  for (int slab = 0; slab < NSLAB; ++slab, ++n) {
    const uint32_t slot = n % Cf::ST, use = n / Cf::ST;
    mbar_wait(rg.full(slot), use & 1);
    if (active) {
      const float* w = reinterpret_cast<const float*>(ring + (size_t)slot * Cf::STAGE) + kg * F + 4 * fg;
      const float* x = xt + (slab * KSLAB + kg) * Cf::XLD + sg * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + i * G * F);
        float xv[8];
        const float4 xa = *reinterpret_cast<const float4*>(x + i * G * Cf::XLD);
        xv[0] = xa.x; xv[1] = xa.y; xv[2] = xa.z; xv[3] = xa.w;
        if constexpr (TS == 6) {
          const float2 xb = *reinterpret_cast<const float2*>(x + i * G * Cf::XLD + 4);
          xv[4] = xb.x; xv[5] = xb.y;
        } else if constexpr (TS == 8) {
          const float4 xb = *reinterpret_cast<const float4*>(x + i * G * Cf::XLD + 4);
          xv[4] = xb.x; xv[5] = xb.y; xv[6] = xb.z; xv[7] = xb.w;
        }
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          acc[0][s] = fmaf(w4.x, xv[s], acc[0][s]);
          acc[1][s] = fmaf(w4.y, xv[s], acc[1][s]);
          acc[2][s] = fmaf(w4.z, xv[s], acc[2][s]);
          acc[3][s] = fmaf(w4.w, xv[s], acc[3][s]);
        }
      }
    }
    // The stage was read through the generic proxy (LDS) and will be overwritten through the async proxy
    // (cp.async.bulk): every reader orders its own reads before the release with a proxy fence.  Without it the
    // refill occasionally overtook a slow warp's loads (measured: 30 % of 24-token runs at B = 256 deviated
    // bitwise; none with the fence - profiles/r02_stack_race.txt).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive_n(rg.empty(slot), 1);  // this warp is done with the stage
  }
#pragma unroll
  for (int off = 1; off < G; off <<= 1)
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int s = 0; s < TS; ++s) acc[f][s] += __shfl_xor_sync(L3_FULL, acc[f][s], off);
}

// RMSNorm (llama3.py:111-114) of the S residual rows, in place on the k-major buffer xt[D][XLD].
template <class Cf>
__device__ __forceinline__ void rms_inplace(float* xt, const float* __restrict__ g, float eps, float* red, float* rinv) {
  constexpr int S = Cf::S, D = Cf::D;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  float ss[S];
#pragma unroll
  for (int s = 0; s < S; ++s) ss[s] = 0.f;
  for (int k = t; k < D; k += Cf::NCOMP) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 q = *reinterpret_cast<const float4*>(xt + k * Cf::XLD + 4 * j);
      v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) ss[s] = fmaf(v[Cf::xslot(s)], v[Cf::xslot(s)], ss[s]);
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
    const float gk = g[k];
#pragma unroll
    for (int sgp = 0; sgp < Cf::SG; ++sgp)
#pragma unroll
      for (int j = 0; j < Cf::TS; ++j) {
        float* p = xt + k * Cf::XLD + sgp * 8 + j;
        *p = *p * rinv[sgp * Cf::TS + j] * gk;
      }
  }
  comp_sync<Cf>();
}

// Partial sums of a row-parallel projection (Wo / Wdown restricted to this CTA's columns) go straight from the
// registers into slot [this rank] of the receive buffer of the CTA that owns the output columns.
template <class Cf, int F, int G>
__device__ __forceinline__ void push_partials(const float (&acc)[4][Cf::TS], uint32_t recv_local, uint32_t pbar_local, int rank) {
  constexpr int NFG = F / 4, NTHR = NFG * Cf::SG * G;
  const int t = threadIdx.x;
  if (t >= NTHR || (t % G) != 0) return;
  const int u = t / G, fg = u % NFG, sg = (u / NFG) % Cf::SG;
  const int f0 = 4 * fg, owner = f0 / Cf::DS, fl = f0 % Cf::DS;  // DS % 4 == 0: the four features share an owner
  const uint32_t dst = mapa(recv_local, (uint32_t)owner) + (uint32_t)(((rank * Cf::DS + fl) * Cf::XLD + sg * 8) * 4);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const uint32_t d = dst + (uint32_t)(f * Cf::XLD * 4);
    st_cluster_f4(d, acc[f][0], acc[f][1], acc[f][2], acc[f][3]);
    if constexpr (Cf::TS == 6) st_cluster_f2(d + 16, acc[f][4], acc[f][5]);
    if constexpr (Cf::TS == 8) st_cluster_f4(d + 16, acc[f][4], acc[f][5], acc[f][6], acc[f][7]);
  }
  mbar_arrive_remote(mapa(pbar_local, (uint32_t)owner));
}

// The owner of residual columns [rank DS, rank DS + DS): x_new = x + sum over ranks (in rank order) of the
// partial sums (llama3.py:253 / :259), kept in xres and written into the k-major residual buffer of EVERY CTA.
template <class Cf>
__device__ __forceinline__ void reduce_and_gather(const float* recv, float* xres, uint32_t xt_local, uint32_t gbar_local, int rank) {
  constexpr int DS = Cf::DS, TS = Cf::TS;
  const int t = threadIdx.x;
  if (t >= DS * Cf::SG) return;
  const int fl = t % DS, sg = t / DS;
  float r[8];
  float* xr = xres + fl * Cf::XLD + sg * 8;
#pragma unroll
  for (int j = 0; j < TS; ++j) r[j] = xr[j];
#pragma unroll
  for (int p = 0; p < Cf::C; ++p) {
    const float* src = recv + (p * DS + fl) * Cf::XLD + sg * 8;
#pragma unroll
    for (int j = 0; j < TS; ++j) r[j] += src[j];
  }
#pragma unroll
  for (int j = 0; j < TS; ++j) xr[j] = r[j];
  const uint32_t off = (uint32_t)((((rank * DS + fl) * Cf::XLD) + sg * 8) * 4);
#pragma unroll
  for (int p = 0; p < Cf::C; ++p) {
    const uint32_t d = mapa(xt_local, (uint32_t)p) + off;
    st_cluster_f4(d, r[0], r[1], r[2], r[3]);
    if constexpr (TS == 6) st_cluster_f2(d + 16, r[4], r[5]);
    if constexpr (TS == 8) st_cluster_f4(d + 16, r[4], r[5], r[6], r[7]);
    mbar_arrive_remote(mapa(gbar_local, (uint32_t)p));
  }
}

// debug dumps, [NL][4][B][D]: kind 0 = q, 1 = attention output (both [S][HD] of this head), 2 / 3 = the residual
// stream after the first / second exchange of the layer (this CTA's DS columns, k-major)
template <class Cf>
__device__ __forceinline__ void dump_cols(const StackArgs& a, int l, int kind, int b0, int s_act, int rank, const float* src_kmajor) {
  const int t = threadIdx.x;
  if (t >= Cf::DS * Cf::SG) return;
  const int fl = t % Cf::DS, sgp = t / Cf::DS;
  for (int j = 0; j < Cf::TS; ++j) {
    const int s = sgp * Cf::TS + j;
    if (s < s_act) a.dbg_x[(((size_t)l * 4 + kind) * a.B + b0 + s) * Cf::D + rank * Cf::DS + fl] = src_kmajor[fl * Cf::XLD + sgp * 8 + j];
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
  constexpr int D = Cf::D, HD = Cf::HD, C = Cf::C, S = Cf::S, TS = Cf::TS, XLD = Cf::XLD, TCH = Cf::TCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint8_t* ring = base;
  float* xt = reinterpret_cast<float*>(base + Cf::OFF_XT);      // [D][XLD]   residual stream, then norm(x), k-major
  float* xres = reinterpret_cast<float*>(base + Cf::OFF_XRES);  // [DS][XLD]  this CTA's columns of the residual stream
  float* q_s = reinterpret_cast<float*>(base + Cf::OFF_Q);      // [S][HD]    rotated q of this head
  float* kn_s = reinterpret_cast<float*>(base + Cf::OFF_KN);    // [S][HD]    this step's rotated k
  float* vn_s = reinterpret_cast<float*>(base + Cf::OFF_VN);    // [S][HD]    this step's v
  float* ctx_t = reinterpret_cast<float*>(base + Cf::OFF_CTX);  // [HD][XLD]  attention output of this head, k-major
  float* h_t = reinterpret_cast<float*>(base + Cf::OFF_H);      // [KE][XLD]  silu(gate) * up of this FFN slice, k-major
  float* recv = reinterpret_cast<float*>(base + Cf::OFF_RECV);  // [C][DS][XLD] partial sums from every rank
  float* part = reinterpret_cast<float*>(base + Cf::OFF_PART);  // [S][NCHMAX][PLD] attention partials
  float* cs = reinterpret_cast<float*>(base + Cf::OFF_CS);      // cos[HD/2] | sin[HD/2] of this position
  float* red = reinterpret_cast<float*>(base + Cf::OFF_RED);
  float* rinv = reinterpret_cast<float*>(base + Cf::OFF_RINV);
  int* tok = reinterpret_cast<int*>(base + Cf::OFF_TOK);
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

  if (t == 0) {
    for (int s = 0; s < Cf::ST; ++s) { mbar_init(rg.full(s), 1); mbar_init(rg.empty(s), Cf::NCW); }
    mbar_init(xb.pbar0, XBars<Cf>::PCOUNT); mbar_init(xb.pbar0 + 8, XBars<Cf>::PCOUNT);
    mbar_init(xb.gbar0, XBars<Cf>::GCOUNT); mbar_init(xb.gbar0 + 8, XBars<Cf>::GCOUNT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist and its shared memory may be written from now on

  if (warp >= Cf::NCW) {
    // ================================================================== producers (one thread per warp)
    // All four walk the same unit sequence; producer p issues the units n with n % NPW == p.  ST % NPW == 0, so
    // the previous use of a unit's stage was issued by the same thread: its wait on `empty` cannot alias.
    if (lane == 0) {
      const int pw = warp - Cf::NCW;
      uint32_t n = 0;
      auto issue = [&](const void* s0, uint32_t bytes0, const void* s1, uint32_t bytes1, uint32_t off1) {
        if ((int)(n % Cf::NPW) == pw) {
          const uint32_t slot = n % Cf::ST, use = n / Cf::ST;
          if (use > 0) mbar_wait(rg.empty(slot), (use - 1) & 1);
          mbar_expect_tx(rg.full(slot), bytes0 + bytes1);
          const uint32_t dst = rg.base + slot * Cf::STAGE;
          bulk_g2s(dst, s0, bytes0, rg.full(slot));
          if (bytes1) bulk_g2s(dst + off1, s1, bytes1, rg.full(slot));
        }
        ++n;
      };
      auto slabs = [&](const float* w, int F, int K, int G) {
        const int ks = 8 * G;
        for (int sl = 0; sl < K / ks; ++sl) issue(w + (size_t)sl * ks * F, (uint32_t)(ks * F * 4), nullptr, 0u, 0u);
      };
      for (int l = 0; l < a.NL; ++l) {
        const StackLayer ly = a.layers[l];
        const float* w = ly.wpack + (size_t)rank * Cf::CTA_FLOATS;
        slabs(w, Cf::FA, Cf::KA, Cf::GA);
        w += Cf::KA * Cf::FA;
        for (int s = 0; s < s_act; ++s) {
          const size_t row0 = ((size_t)(b0 + s) * C + rank) * a.M;
          for (int c = 0; c < nch; ++c) {
            const uint32_t rows = (uint32_t)min(TCH, pos - c * TCH);
            issue(ly.ck + (row0 + (size_t)c * TCH) * HD, rows * HD * 4, ly.cv + (row0 + (size_t)c * TCH) * HD, rows * HD * 4,
                  (uint32_t)(TCH * HD * 4));
          }
        }
        slabs(w, Cf::FB, Cf::KB, Cf::GB);
        w += Cf::KB * Cf::FB;
        slabs(w, Cf::FC, Cf::KC, Cf::GC);
        w += Cf::KC * Cf::FC;
        slabs(w, Cf::FE, Cf::KE, Cf::GE);
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
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (tok[s] >= 0) v[Cf::xslot(s)] = a.embed[(size_t)tok[s] * D + k];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(xt + k * XLD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      if (k >= rank * Cf::DS && k < (rank + 1) * Cf::DS) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(xres + (k - rank * Cf::DS) * XLD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    comp_sync<Cf>();
    SK_STAMP(a, 0);

    float acc[4][TS];
    for (int l = 0; l < a.NL; ++l) {
      const StackLayer ly = a.layers[l];
      // ---- q, k, v = rope(norm(x) Wqkv^T) of head `rank`; k, v -> cache          llama3.py:248, 166-187
      rms_inplace<Cf>(xt, ly.norm_in, a.eps, red, rinv);
      gemm_phase<Cf, Cf::FA, Cf::KA, Cf::GA>(rg, ring, xt, n, acc);
      {
        constexpr int NFG = Cf::FA / 4, G = Cf::GA;
        if (t < NFG * Cf::SG * G && (t % G) == 0) {
          const int u = t / G, fg = u % NFG, sg = (u / NFG) % Cf::SG;
          const int f0 = 4 * fg, region = f0 / HD, d0 = f0 % HD;  // 0 q, 1 k, 2 v
          float c0 = 1.f, s0 = 0.f, c1 = 1.f, s1 = 0.f;
          if (region < 2) { c0 = cs[d0 >> 1]; s0 = cs[HD / 2 + (d0 >> 1)]; c1 = cs[(d0 >> 1) + 1]; s1 = cs[HD / 2 + (d0 >> 1) + 1]; }
          float* dst_s = region == 0 ? q_s : (region == 1 ? kn_s : vn_s);
          float* cache = region == 1 ? ly.ck : ly.cv;
#pragma unroll
          for (int j = 0; j < TS; ++j) {
            const int s = sg * TS + j;
            // interleaved-pair rotation (llama3.py:41-76); v passes through with (cos, sin) = (1, 0)
            const float4 o = make_float4(acc[0][j] * c0 - acc[1][j] * s0, acc[0][j] * s0 + acc[1][j] * c0,
                                         acc[2][j] * c1 - acc[3][j] * s1, acc[2][j] * s1 + acc[3][j] * c1);
            *reinterpret_cast<float4*>(dst_s + s * HD + d0) = o;
            if (region > 0 && s < s_act)  // cache append at this position (llama3.py:184-185)
              *reinterpret_cast<float4*>(cache + (((size_t)(b0 + s) * C + rank) * a.M + pos) * HD + d0) = o;
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
            mbar_wait(rg.full(slot), use & 1);
            const float* Ks = reinterpret_cast<const float*>(ring + (size_t)slot * Cf::STAGE);
            const float* Vs = Ks + TCH * HD;
            constexpr int NP = TCH / 8;
            float sc[NP];
            float mx = -INFINITY;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const int r = p * 8 + sub;
              float d = 0.f;
              if (r < rows) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                  const float4 k4 = *reinterpret_cast<const float4*>(Ks + r * HD + (sl + 4 * i) * 4);
                  d = fmaf(q4[i].x, k4.x, d); d = fmaf(q4[i].y, k4.y, d); d = fmaf(q4[i].z, k4.z, d); d = fmaf(q4[i].w, k4.w, d);
                }
              }
              d += __shfl_xor_sync(L3_FULL, d, 1);
              d += __shfl_xor_sync(L3_FULL, d, 2);
              sc[p] = r < rows ? d * scale : -INFINITY;
              mx = fmaxf(mx, sc[p]);
            }
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 8));
            mx = fmaxf(mx, __shfl_xor_sync(L3_FULL, mx, 16));  // rows >= 1: finite
            float o[4 * CPL], lsum = 0.f;
#pragma unroll
            for (int e = 0; e < 4 * CPL; ++e) o[e] = 0.f;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const int r = p * 8 + sub;
              if (r < rows) {
                const float pw_ = expf(sc[p] - mx);
                lsum += pw_;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                  const float4 v4 = *reinterpret_cast<const float4*>(Vs + r * HD + (sl + 4 * i) * 4);
                  o[4 * i] = fmaf(pw_, v4.x, o[4 * i]); o[4 * i + 1] = fmaf(pw_, v4.y, o[4 * i + 1]);
                  o[4 * i + 2] = fmaf(pw_, v4.z, o[4 * i + 2]); o[4 * i + 3] = fmaf(pw_, v4.w, o[4 * i + 3]);
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
            if (sub == 0) {
              float* pp = part + (s * Cf::NCHMAX + c) * Cf::PLD;
#pragma unroll
              for (int i = 0; i < CPL; ++i)
                *reinterpret_cast<float4*>(pp + (sl + 4 * i) * 4) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
              if (sl == 0) { pp[HD] = mx; pp[HD + 1] = lsum; }
            }
          }
        }
        comp_sync<Cf>();
        // merge: 4 lanes per sequence (warps 0 .. ceil(4 S / 32) - 1 take part as whole warps)
        if (warp < (4 * S + 31) / 32) {
          constexpr int CPL = HD / 16;
          const int s = min(t >> 2, S - 1), sl = t & 3;
          const int nc = s < s_act ? nch : 0;
          float4 q4[CPL], kv4[CPL];
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            q4[i] = *reinterpret_cast<const float4*>(q_s + s * HD + (sl + 4 * i) * 4);
            kv4[i] = *reinterpret_cast<const float4*>(kn_s + s * HD + (sl + 4 * i) * 4);
            d = fmaf(q4[i].x, kv4[i].x, d); d = fmaf(q4[i].y, kv4[i].y, d); d = fmaf(q4[i].z, kv4[i].z, d); d = fmaf(q4[i].w, kv4[i].w, d);
          }
          d += __shfl_xor_sync(L3_FULL, d, 1);
          d += __shfl_xor_sync(L3_FULL, d, 2);
          const float s_new = d * scale;
          float mx = s_new;
          const float* pp = part + s * Cf::NCHMAX * Cf::PLD;
          for (int c = 0; c < nc; ++c) mx = fmaxf(mx, pp[c * Cf::PLD + HD]);
          const float wn = expf(s_new - mx);
          float lsum = wn, o[4 * CPL];
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            const float4 v4 = *reinterpret_cast<const float4*>(vn_s + s * HD + (sl + 4 * i) * 4);
            o[4 * i] = wn * v4.x; o[4 * i + 1] = wn * v4.y; o[4 * i + 2] = wn * v4.z; o[4 * i + 3] = wn * v4.w;
          }
          for (int c = 0; c < nc; ++c) {
            const float w = expf(pp[c * Cf::PLD + HD] - mx);
            lsum = fmaf(pp[c * Cf::PLD + HD + 1], w, lsum);
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
              const float4 p4 = *reinterpret_cast<const float4*>(pp + c * Cf::PLD + (sl + 4 * i) * 4);
              o[4 * i] = fmaf(p4.x, w, o[4 * i]); o[4 * i + 1] = fmaf(p4.y, w, o[4 * i + 1]);
              o[4 * i + 2] = fmaf(p4.z, w, o[4 * i + 2]); o[4 * i + 3] = fmaf(p4.w, w, o[4 * i + 3]);
            }
          }
          if ((t >> 2) < S) {
            const float inv = 1.0f / lsum;
            const int xs = Cf::xslot(s);
#pragma unroll
            for (int i = 0; i < CPL; ++i)
#pragma unroll
              for (int e = 0; e < 4; ++e) ctx_t[((sl + 4 * i) * 4 + e) * XLD + xs] = o[4 * i + e] * inv;
          }
        }
        comp_sync<Cf>();
      }
      if (a.dbg_x) dump_cols<Cf>(a, l, 1, b0, s_act, rank, ctx_t);  // HD == DS at this shape
      if (l < 6) SK_STAMP(a, 2 + l * 10);
      // ---- x += ctx Wo^T                                                        llama3.py:210-211, 253
      gemm_phase<Cf, Cf::FB, Cf::KB, Cf::GB>(rg, ring, ctx_t, n, acc);
      push_partials<Cf, Cf::FB, Cf::GB>(acc, smem_u32(recv), xb.pbar0 + 8 * (xb.np & 1), rank);
      if (l < 6) SK_STAMP(a, 3 + l * 10);
      mbar_wait_cluster(xb.pbar0 + 8 * (xb.np & 1), (xb.np >> 1) & 1);  // all partial sums for my columns are here
      xb.np += 1;
      reduce_and_gather<Cf>(recv, xres, smem_u32(xt), xb.gbar0 + 8 * (xb.ng & 1), rank);
      mbar_wait_cluster(xb.gbar0 + 8 * (xb.ng & 1), (xb.ng >> 1) & 1);  // the whole new residual stream is here
      xb.ng += 1;
      if (a.dbg_x) dump_cols<Cf>(a, l, 2, b0, s_act, rank, xres);
      if (l < 6) SK_STAMP(a, 4 + l * 10);
      // ---- h = silu(norm(x) Wgate^T) * (norm(x) Wup^T) of FFN slice `rank`          llama3.py:256, 99-101
      rms_inplace<Cf>(xt, ly.norm_post, a.eps, red, rinv);
      gemm_phase<Cf, Cf::FC, Cf::KC, Cf::GC>(rg, ring, xt, n, acc);
      {
        constexpr int NFG = Cf::FC / 4, G = Cf::GC;
        if (t < NFG * Cf::SG * G && (t % G) == 0) {
          const int u = t / G, fg = u % NFG, sg = (u / NFG) % Cf::SG;
          // features 4 fg .. 4 fg + 3 = (gate, up) of h columns 2 fg and 2 fg + 1 of this slice
#pragma unroll
          for (int j = 0; j < TS; ++j) {
            h_t[(2 * fg) * XLD + sg * 8 + j] = silu_ref(acc[0][j]) * acc[1][j];
            h_t[(2 * fg + 1) * XLD + sg * 8 + j] = silu_ref(acc[2][j]) * acc[3][j];
          }
        }
      }
      comp_sync<Cf>();
      if (l < 6) SK_STAMP(a, 5 + l * 10);
      // ---- x += h Wdown^T                                                        llama3.py:102, 259
      gemm_phase<Cf, Cf::FE, Cf::KE, Cf::GE>(rg, ring, h_t, n, acc);
      push_partials<Cf, Cf::FE, Cf::GE>(acc, smem_u32(recv), xb.pbar0 + 8 * (xb.np & 1), rank);
      if (l < 6) SK_STAMP(a, 6 + l * 10);
      mbar_wait_cluster(xb.pbar0 + 8 * (xb.np & 1), (xb.np >> 1) & 1);  // all partial sums for my columns are here
      xb.np += 1;
      reduce_and_gather<Cf>(recv, xres, smem_u32(xt), xb.gbar0 + 8 * (xb.ng & 1), rank);
      mbar_wait_cluster(xb.gbar0 + 8 * (xb.ng & 1), (xb.ng >> 1) & 1);  // the whole new residual stream is here
      xb.ng += 1;
      if (l < 6) SK_STAMP(a, 7 + l * 10);
      if (a.dbg_x) dump_cols<Cf>(a, l, 3, b0, s_act, rank, xres);
    }
    // ---- final norm (llama3.py:304); every CTA writes its DS columns of the LM head's operand rows
    rms_inplace<Cf>(xt, a.norm_final, a.eps, red, rinv);
    for (int i = t; i < s_act * (Cf::DS / 4); i += Cf::NCOMP) {
      const int s = i / (Cf::DS / 4), k4 = (i % (Cf::DS / 4)) * 4 + rank * Cf::DS;
      const int xs = Cf::xslot(s);
      float4 hi, lo;
      split_tf32(xt[(k4 + 0) * XLD + xs], hi.x, lo.x);
      split_tf32(xt[(k4 + 1) * XLD + xs], hi.y, lo.y);
      split_tf32(xt[(k4 + 2) * XLD + xs], hi.z, lo.z);
      split_tf32(xt[(k4 + 3) * XLD + xs], hi.w, lo.w);
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
