// Decode attention (L == 1) as device functions shared by the standalone kernel (attention.cu)
// and the persistent decode kernel (decode_mega.cu).
//
// One work item = (split, kv-head group, sequence).  Each group of LPK lanes owns a key at a
// time, holding EPL = HD / LPK dimensions (as interleaved 16-byte chunks), and keeps its own online-softmax state (m, l, o) for
// the NREP query heads of the group, so K and V are read exactly once for all heads sharing
// them (repeat_kv, llama3.py:79-83, is index math).  Lane groups are merged through shared
// memory; with nsplit > 1 items emit (m, l, o) partials and the last item of a head group to
// arrive merges them (flash-decoding without a second launch).
#pragma once
#include "common.cuh"

// A cache row (HD elements) is NCH chunks of 16 bytes.  LPK lanes share a key; lane sl owns the
// chunks sl, sl + LPK, sl + 2 LPK, ... so that every load instruction of the lane group covers
// LPK * 16 contiguous bytes (whole 32-byte sectors, no partial-sector refetch for head_dim 48 / 96).
template <int HD, typename KVT> struct DecodeCfg {
  static constexpr int VEC = 16 / (int)sizeof(KVT);   // elements per 16-byte chunk
  static constexpr int NCH = HD / VEC;
  static constexpr int LPK = (NCH % 16 == 0) ? 16 : (NCH % 8 == 0) ? 8 : (NCH % 4 == 0) ? 4 : (NCH % 2 == 0) ? 2 : 1;
  static constexpr int CPL = NCH / LPK;                // chunks per lane
  static constexpr int EPL = CPL * VEC;                // elements per lane
  static constexpr int KPW = 32 / LPK;                 // keys per warp pass
  static_assert(HD % VEC == 0, "head_dim must be a multiple of the 16-byte chunk");
  // element e of the lane (0 <= e < EPL) is dimension dim(sl, e) of the head
  __device__ static __forceinline__ int dim(int sl, int e) { return ((e / VEC) * LPK + sl) * VEC + (e % VEC); }
};

// Where a cache row is read from.  KV_STREAM: global memory through the non-coherent streaming path.  KV_COHERENT: the
// row may have been written earlier in the SAME launch by another SM (persistent kernel): read through L2
// (ld.global.cg).  KV_STAGED: a copy of the row staged in shared memory.
enum { KV_STREAM = 0, KV_COHERENT = 1, KV_STAGED = 2 };
template <int SRC> __device__ __forceinline__ uint4 kv_ld16(const void* p) {
  if constexpr (SRC == KV_COHERENT) return __ldcg(reinterpret_cast<const uint4*>(p));
  else if constexpr (SRC == KV_STAGED) return *reinterpret_cast<const uint4*>(p);
  else return ldg_stream16(p);
}

// the lane's CPL chunks of one cache row (row points at the row start)
template <int HD, int COH>
__device__ __forceinline__ void load_row(const float* row, int sl, float (&v)[DecodeCfg<HD, float>::EPL]) {
  using C = DecodeCfg<HD, float>;
#pragma unroll
  for (int c = 0; c < C::CPL; ++c) {
    const uint4 r = kv_ld16<COH>(row + (c * C::LPK + sl) * 4);
    v[4 * c] = __uint_as_float(r.x); v[4 * c + 1] = __uint_as_float(r.y);
    v[4 * c + 2] = __uint_as_float(r.z); v[4 * c + 3] = __uint_as_float(r.w);
  }
}
template <int HD, int COH>
__device__ __forceinline__ void load_row(const bf16* row, int sl, float (&v)[DecodeCfg<HD, bf16>::EPL]) {
  using C = DecodeCfg<HD, bf16>;
#pragma unroll
  for (int c = 0; c < C::CPL; ++c) {
    const uint4 r = kv_ld16<COH>(row + (c * C::LPK + sl) * 8);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[8 * c + 2 * j] = __uint_as_float(w[j] << 16);
      v[8 * c + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}

__device__ __forceinline__ void st_tagged(unsigned long long* p, float v, unsigned tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}

template <int HD, int NREP, int NW, typename KVT> struct AttnDecodeSmem {
  static constexpr int NSLOT = NW;  // one merged state per warp
  float m[NREP][NSLOT];
  float l[NREP][NSLOT];
  float o[NREP][NSLOT][HD];
  float cw[NREP][32];
  float cl[NREP];
  int last;
};

// Flash-decoding merge of the nsplit (<= 32) partial results of NREP heads by the NW warps of
// a work group: the (m, l) pairs are fetched in parallel, lane s of a warp turns split s into
// its weight exp(m_s - max), and every output element then sums nsplit independent L2 loads.
template <int HD, int NREP, int NW, typename Sync>
__device__ __forceinline__ void combine_splits(const AttnArgs& a, int b, int head0, int tid, float (*cmb_w)[32],
                                               float* cmb_l, Sync sync) {
  const int lane = tid & 31, warp = tid >> 5;
  for (int r = warp; r < NREP; r += NW) {
    const size_t q0 = ((size_t)b * a.HN + head0 + r) * a.nsplit;
    float ms = -INFINITY, ls = 0.f;
    if (lane < a.nsplit) {
      const float2 ml = __ldcg(reinterpret_cast<const float2*>(a.part_ml + (q0 + lane) * 2));
      ms = ml.x; ls = ml.y;
    }
    const float mx = warp_max(ms);
    const float w = ms > -INFINITY ? expf(ms - mx) : 0.f;  // empty split -> weight 0
    const float lsum = warp_sum(ls * w);
    cmb_w[r][lane] = w;
    if (lane == 0) cmb_l[r] = lsum;
  }
  sync();
  for (int idx = tid; idx < NREP * HD; idx += NW * 32) {
    const int r = idx / HD, d = idx % HD, head = head0 + r;
    const float* po = a.part_o + ((size_t)b * a.HN + head) * a.nsplit * HD + d;
    float osum = 0.f;
#pragma unroll 8
    for (int s = 0; s < a.nsplit; ++s) osum = fmaf(__ldcg(po + (size_t)s * HD), cmb_w[r][s], osum);
    const float v = osum / cmb_l[r];
    const size_t oi = ((size_t)b * a.HN + head) * HD + d;
    if (a.out_lo) { float hi, lo; split_tf32(v, hi, lo); __stcg(a.out + oi, hi); __stcg(a.out_lo + oi, lo); }
    else if (a.out) __stcg(a.out + oi, v);
    if (a.out_bf16) a.out_bf16[oi] = __float2bfloat16_rn(v);
    if (a.out_ll) st_tagged(a.out_ll + oi, v, a.out_tag);
  }
}

// The lane's dimensions of the NREP query rows (already rotated, fp32) and an empty online-softmax state.
template <int HD, int NREP, typename KVT>
__device__ __forceinline__ void attn_decode_init(const AttnArgs& a, int b, int head0, int sl,
                                                 float (&q)[NREP][DecodeCfg<HD, KVT>::EPL],
                                                 float (&o)[NREP][DecodeCfg<HD, KVT>::EPL], float (&m)[NREP], float (&l)[NREP]) {
  using C = DecodeCfg<HD, KVT>;
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    const float* qp = a.q + ((size_t)b * a.HN + head0 + r) * HD;
#pragma unroll
    for (int e = 0; e < C::EPL; e += 2) {
      float2 t = __ldcg(reinterpret_cast<const float2*>(qp + C::dim(sl, e)));
      q[r][e] = t.x; q[r][e + 1] = t.y;
    }
#pragma unroll
    for (int e = 0; e < C::EPL; ++e) o[r][e] = 0.f;
    m[r] = -INFINITY;
    l[r] = 0.f;
  }
}

// One batch of U keys per lane group (kk / vv = the lane's dimensions of the keys' K and V rows, ok = key exists) folded
// into the lane group's online-softmax state of the NREP query heads.
template <int NREP, int EPL, int LPK, int U>
__device__ __forceinline__ void attn_decode_batch(const float (&q)[NREP][EPL], const float (&kk)[U][EPL],
                                                  const float (&vv)[U][EPL], const bool (&ok)[U], float scale,
                                                  float (&o)[NREP][EPL], float (&m)[NREP], float (&l)[NREP]) {
#pragma unroll
  for (int r = 0; r < NREP; ++r) {
    float s[U];
    float mx = m[r];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) d = fmaf(q[r][e], kk[u][e], d);
#pragma unroll
      for (int off = LPK / 2; off > 0; off >>= 1) d += __shfl_xor_sync(L3_FULL, d, off);
      s[u] = ok[u] ? d * scale : -INFINITY;
      mx = fmaxf(mx, s[u]);
    }
    if (mx > -INFINITY) {
      const float alpha = expf(m[r] - mx);  // m = -inf -> 0
      float ps = 0.f;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[r][e] *= alpha;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float p = expf(s[u] - mx);  // s = -inf -> 0
        ps += p;
#pragma unroll
        for (int e = 0; e < EPL; ++e) o[r][e] = fmaf(p, vv[u][e], o[r][e]);
      }
      l[r] = l[r] * alpha + ps;
      m[r] = mx;
    }
  }
}

// Second half of the end of a work item: every warp's merged state (m, l, o) of the NREP heads is in shared memory;
// merges the NW warps, then writes the result (nsplit == 1) or publishes the item's partials, the last item of a
// (sequence, head group) to arrive combining all of them.
template <int HD, int NREP, typename KVT, int NW, typename Sync>
__device__ __forceinline__ void attn_decode_finish_smem(const AttnArgs& a, int split, int grp, int ngrp, int b, int tid,
                                                        AttnDecodeSmem<HD, NREP, NW, KVT>& sm, Sync sync) {
  constexpr int NSLOT = NW;
  const int head0 = grp * NREP;
  sync();
  for (int idx = tid; idx < NREP * HD; idx += NW * 32) {
    const int r = idx / HD, d = idx % HD;
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) mx = fmaxf(mx, sm.m[r][s]);
    float lsum = 0.f, osum = 0.f;
    if (mx > -INFINITY) {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const float w = expf(sm.m[r][s] - mx);
        lsum = fmaf(sm.l[r][s], w, lsum);
        osum = fmaf(sm.o[r][s][d], w, osum);
      }
    }
    const int head = head0 + r;
    if (a.nsplit == 1) {
      const float v = osum / lsum;
      const size_t oi = ((size_t)b * a.HN + head) * HD + d;
      if (a.out_lo) { float hi, lo; split_tf32(v, hi, lo); __stcg(a.out + oi, hi); __stcg(a.out_lo + oi, lo); }
      else if (a.out) __stcg(a.out + oi, v);
      if (a.out_bf16) a.out_bf16[oi] = __float2bfloat16_rn(v);
      if (a.out_ll) st_tagged(a.out_ll + oi, v, a.out_tag);
    } else {
      const size_t pi = ((size_t)b * a.HN + head) * a.nsplit + split;
      __stcg(a.part_o + pi * HD + d, osum);
      if (d == 0) { __stcg(a.part_ml + pi * 2, mx); __stcg(a.part_ml + pi * 2 + 1, lsum); }
    }
  }
  if (a.nsplit > 1 && a.counters) {
    // the last item of this (sequence, head group) to publish its partials combines all of them: every thread fences
    // its own partial stores, the barrier collects them, one thread counts; the last arriver fences again before it
    // reads the other items' partials
    __threadfence();
    sync();
    if (tid == 0) {
      int* cnt = a.counters + (size_t)b * ngrp + grp;
      const int old = atomicAdd(cnt, 1);
      sm.last = (old == a.nsplit - 1);
      if (sm.last) *cnt = 0;  // ready for the next launch
    }
    sync();
    if (sm.last) {
      __threadfence();
      combine_splits<HD, NREP, NW>(a, b, head0, tid, sm.cw, &sm.cl[0], sync);
    }
  }
  sync();  // the shared state may be reused by the caller's next item
}

// The end of a work item: merges the KPW lane groups of each warp with shuffles and the NW warps through shared memory,
// then writes the result (nsplit == 1) or publishes the item's (m, l, o) partials, the last item of a (sequence, head
// group) to arrive combining all of them.
template <int HD, int NREP, typename KVT, int NW, typename Sync>
__device__ __forceinline__ void attn_decode_finish(const AttnArgs& a, int split, int grp, int ngrp, int b, int tid,
                                                   float (&o)[NREP][DecodeCfg<HD, KVT>::EPL], float (&m)[NREP],
                                                   float (&l)[NREP], AttnDecodeSmem<HD, NREP, NW, KVT>& sm, Sync sync) {
  using C = DecodeCfg<HD, KVT>;
  constexpr int LPK = C::LPK, EPL = C::EPL;
  const int lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPK, sl = lane % LPK;
#pragma unroll
  for (int off = LPK; off < 32; off <<= 1) {
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      const float mo = __shfl_xor_sync(L3_FULL, m[r], off), lo = __shfl_xor_sync(L3_FULL, l[r], off);
      const float mn = fmaxf(m[r], mo);
      const float wa = m[r] > -INFINITY ? expf(m[r] - mn) : 0.f;
      const float wb = mo > -INFINITY ? expf(mo - mn) : 0.f;
      l[r] = l[r] * wa + lo * wb;
#pragma unroll
      for (int e = 0; e < EPL; ++e) o[r][e] = o[r][e] * wa + __shfl_xor_sync(L3_FULL, o[r][e], off) * wb;
      m[r] = mn;
    }
  }
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < NREP; ++r) {
      if (sl == 0) { sm.m[r][warp] = m[r]; sm.l[r][warp] = l[r]; }
#pragma unroll
      for (int e = 0; e < EPL; ++e) sm.o[r][warp][C::dim(sl, e)] = o[r][e];
    }
  }
  attn_decode_finish_smem<HD, NREP, KVT, NW>(a, split, grp, ngrp, b, tid, sm, sync);
}

// tid in [0, NW * 32); sync() is a barrier over exactly those threads; ngrp = head groups per
// sequence (the counter index space); T = keys visible to the query (start_pos + 1).
// U = key batches in flight per lane group (loads of U * KPW * NW keys are issued before any is used)
template <int HD, int NREP, typename KVT, int NW, bool COH, typename Sync, int U = 2>
__device__ __forceinline__ void attn_decode_item(const AttnArgs& a, int nrep_actual, int split, int grp, int ngrp, int b,
                                                 int T, int tid, AttnDecodeSmem<HD, NREP, NW, KVT>& sm, Sync sync) {
  using C = DecodeCfg<HD, KVT>;
  constexpr int LPK = C::LPK, EPL = C::EPL, KPW = C::KPW;
  const int head0 = grp * NREP;           // first query head of this item
  const int kvh = head0 / nrep_actual;    // its kv head (llama3.py:79-83)
  const int chunk = (T + a.nsplit - 1) / a.nsplit;
  const int t0 = split * chunk;
  const int t1 = min(T, t0 + chunk);
  const int lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPK, sl = lane % LPK;
  const float scale = 1.0f / sqrtf((float)HD);

  float q[NREP][EPL], o[NREP][EPL], m[NREP], l[NREP];
  attn_decode_init<HD, NREP, KVT>(a, b, head0, sl, q, o, m, l);

  const KVT* kbase = (const KVT*)a.cache_k + ((size_t)b * a.KVHN + kvh) * a.M * HD;
  const KVT* vbase = (const KVT*)a.cache_v + ((size_t)b * a.KVHN + kvh) * a.M * HD;
  constexpr int KSTRIDE = NW * KPW;
  for (int base = t0 + warp * KPW; base < t1; base += KSTRIDE * U) {
    float kk[U][EPL], vv[U][EPL];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = base + sub + u * KSTRIDE;
      ok[u] = t < t1;
      if (ok[u]) {
        load_row<HD, COH>(kbase + (size_t)t * HD, sl, kk[u]);
        load_row<HD, COH>(vbase + (size_t)t * HD, sl, vv[u]);
      } else {
#pragma unroll
        for (int e = 0; e < EPL; ++e) { kk[u][e] = 0.f; vv[u][e] = 0.f; }
      }
    }
    attn_decode_batch<NREP, EPL, LPK, U>(q, kk, vv, ok, scale, o, m, l);
  }

  attn_decode_finish<HD, NREP, KVT, NW>(a, split, grp, ngrp, b, tid, o, m, l, sm, sync);
}
