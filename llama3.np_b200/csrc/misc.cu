// Small kernels around the projections: embedding gather (llama3.py:287), stand-alone RMSNorm
// (llama3.py:111-114, used ahead of the tiled GEMMs), greedy argmax (llama3.py:320), weight
// packing / random init, cache layout conversion for state inspection, and the per-op
// RoPE / SwiGLU entry points used by the parity tests.
#include <algorithm>

#include "common.cuh"

// -------------------------------------------------------------------------- embedding gather
template <typename WT>
__global__ void embed_kernel(const WT* __restrict__ table, const int32_t* __restrict__ ids, int ids_ld, int ids_off,
                             int L, int rows, int D, int vocab, float* __restrict__ x) {
  const int r = blockIdx.x;
  pdl_launch();
  pdl_wait();
  if (r >= rows) return;
  const int b = r / L, t = r - b * L;
  // host entry points validate ids; device-resident ids (l3_forward_dev, generate loops) are clamped here so that a
  // bad id reads a wrong row instead of foreign memory
  const int id = min(max(ids[(size_t)b * ids_ld + ids_off + t], 0), vocab - 1);
  const WT* src = table + (size_t)id * D;
  float* dst = x + (size_t)r * D;
  for (int k = threadIdx.x * 4; k < D; k += blockDim.x * 4) {
    float4 o;
    o.x = to_f32(src[k]); o.y = to_f32(src[k + 1]); o.z = to_f32(src[k + 2]); o.w = to_f32(src[k + 3]);
    *reinterpret_cast<float4*>(dst + k) = o;
  }
}

cudaError_t launch_embed(const void* table, bool bf16_table, const int32_t* ids, int ids_ld, int ids_off, int L,
                         int rows, int D, int vocab, float* x, cudaStream_t s) {
  const int threads = D >= 1024 ? 256 : 64;
  if (bf16_table)
    return launch_k(embed_kernel<bf16>, dim3(rows), dim3(threads), 0, s, (const bf16*)table, ids, ids_ld, ids_off, L, rows, D, vocab, x);
  return launch_k(embed_kernel<float>, dim3(rows), dim3(threads), 0, s, (const float*)table, ids, ids_ld, ids_off, L, rows, D, vocab, x);
}

// -------------------------------------------------------------------------- RMSNorm
// One CTA per row.  Rows of up to 32 values per thread (D <= 8192 at 256 threads) are read ONCE into registers together
// with the norm weights - one round trip to L2 instead of three dependent ones (row, row again, weights), which is
// what a 6 us launch at 32 rows x 4096 consisted of; longer rows are read twice (second read hits L1).  Emits the
// operand form its consumer wants: fp32, bf16, or the exact TF32 (hi, lo) pair.
__device__ __forceinline__ void rmsnorm_emit(float4 v, float4 g, float rinv, size_t o, float* __restrict__ out,
                                             bf16* __restrict__ out_bf16, float* __restrict__ out_lo) {
  v.x = v.x * rinv * g.x; v.y = v.y * rinv * g.y; v.z = v.z * rinv * g.z; v.w = v.w * rinv * g.w;
  if (out_lo) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(out + o) = hi;
    *reinterpret_cast<float4*>(out_lo + o) = lo;
  } else if (out) *reinterpret_cast<float4*>(out + o) = v;
  if (out_bf16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out_bf16 + o) = pk;
  }
}

__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      float eps, int rows, int D, int src_mul, int src_add,
                                                      float* __restrict__ out, bf16* __restrict__ out_bf16,
                                                      float* __restrict__ out_lo, const int32_t* __restrict__ src_rows) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, nw = nt >> 5;
  pdl_launch();
  float4 gv[8];  // the norm weights do not depend on the previous kernel
  const bool cached = D <= nt * 32;
  if (cached) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = (tid + i * nt) * 4;
      gv[i] = k < D ? __ldg(reinterpret_cast<const float4*>(w + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  pdl_wait();
  const float* src = x + (src_rows ? (size_t)src_rows[r] : (size_t)r * src_mul + src_add) * D;
  float4 xv[8];
  float ss = 0.f;
  if (cached) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = (tid + i * nt) * 4;
      xv[i] = k < D ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // the same order of additions as the two-pass form below: element blocks in ascending k per thread
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
  } else {
    for (int k = tid * 4; k < D; k += nt * 4) {
      float4 v = *reinterpret_cast<const float4*>(src + k);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  ss = 0.f;
  for (int i = 0; i < nw; ++i) ss += red[i];
  const float rinv = 1.0f / sqrtf(ss / (float)D + eps);
  if (cached) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = (tid + i * nt) * 4;
      if (k < D) rmsnorm_emit(xv[i], gv[i], rinv, (size_t)r * D + k, out, out_bf16, out_lo);
    }
  } else {
    for (int k = tid * 4; k < D; k += nt * 4)
      rmsnorm_emit(*reinterpret_cast<const float4*>(src + k), *reinterpret_cast<const float4*>(w + k), rinv, (size_t)r * D + k,
                   out, out_bf16, out_lo);
  }
}

cudaError_t launch_rmsnorm(const float* x, const float* w, float eps, int rows, int D, int src_mul, int src_add,
                           float* out, bf16* out_bf16, float* out_lo, cudaStream_t s, const int32_t* src_rows) {
  // 128 threads as before wherever that was the whole story (the sum of squares keeps its order of additions: the
  // fp32 path is token-identical to the oracle); 256 threads for rows that would not fit 128 threads' registers
  const int threads = D > 128 * 32 ? 256 : 128;
  return launch_k(rmsnorm_kernel, dim3(rows), dim3(threads), 0, s, x, w, eps, rows, D, src_mul, src_add, out, out_bf16,
                  out_lo, src_rows);
}

// -------------------------------------------------------------------------- greedy argmax
// One CTA per row; first maximum wins (NumPy argmax semantics, llama3.py:320).
__device__ __forceinline__ void amax_merge(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ logits, int n,
                                                      int32_t* __restrict__ next_ids, int64_t* __restrict__ out64,
                                                      int out_stride, const int* __restrict__ step_ptr,
                                                      unsigned long long* __restrict__ best_keys, int col_offset) {
  __shared__ float sv[32];
  __shared__ int si[32];
  pdl_launch();
  pdl_wait();
  const float* row = logits + (size_t)blockIdx.x * n;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = row[i];
    if (v > best) { best = v; bi = i; }  // ascending i per thread: strict > keeps the first
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(L3_FULL, best, o);
    int oi = __shfl_xor_sync(L3_FULL, bi, o);
    amax_merge(best, bi, ov, oi);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(L3_FULL, best, o);
      int oi = __shfl_xor_sync(L3_FULL, bi, o);
      amax_merge(best, bi, ov, oi);
    }
    if (lane == 0 && best_keys) {  // vocabulary-sharded LM head: emit the mergeable (value, index) key
      best_keys[blockIdx.x] = bi == 0x7fffffff ? 0ull : argmax_key(best, col_offset + bi);
    } else if (lane == 0) {
      if (bi == 0x7fffffff) bi = 0;  // all -inf / NaN row: NumPy would return 0 for all -inf
      if (next_ids) next_ids[blockIdx.x] = bi;
      if (out64) out64[(size_t)blockIdx.x * out_stride + (step_ptr ? *step_ptr : 0)] = (int64_t)bi;
    }
  }
}

cudaError_t launch_argmax(const float* logits, int rows, int n, int32_t* next_ids, int64_t* out64, int out_stride,
                          const int* step_ptr, cudaStream_t s) {
  const int threads = n >= 65536 ? 1024 : (n >= 8192 ? 512 : 128);
  return launch_k(argmax_kernel, dim3(rows), dim3(threads), 0, s, logits, n, next_ids, out64, out_stride, step_ptr,
                  (unsigned long long*)nullptr, 0);
}

cudaError_t launch_argmax_keys(const float* logits, int rows, int n, int col_offset, unsigned long long* best,
                               cudaStream_t s) {
  const int threads = n >= 65536 ? 1024 : (n >= 8192 ? 512 : 128);
  return launch_k(argmax_kernel, dim3(rows), dim3(threads), 0, s, logits, n, (int32_t*)nullptr, (int64_t*)nullptr, 0,
                  (const int*)nullptr, best, col_offset);
}

// [G, rows, n] rank-major all-gathered slices -> [rows, G * n]
__global__ void gather_permute_kernel(const float* __restrict__ in, int G, int rows, int n, float* __restrict__ out) {
  const int64_t total = (int64_t)G * rows * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n);
    const int r = (int)((i / n) % rows);
    const int g = (int)(i / ((int64_t)n * rows));
    out[((int64_t)r * G + g) * n + c] = in[i];
  }
}
cudaError_t launch_gather_permute(const float* in, int G, int rows, int n, float* out, cudaStream_t s) {
  const int64_t total = (int64_t)G * rows * n;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  return launch_k(gather_permute_kernel, dim3(grid), dim3(256), 0, s, in, G, rows, n, out);
}

__global__ void argmax_finalize_kernel(unsigned long long* __restrict__ best, int rows, int32_t* __restrict__ next_ids,
                                       int64_t* __restrict__ out64, int out_stride, const int* __restrict__ step_ptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch();
  pdl_wait();
  if (r >= rows) return;
  const unsigned long long k = best[r];
  best[r] = 0ull;
  const int idx = k ? (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull)) : 0;
  if (next_ids) next_ids[r] = idx;
  if (out64) out64[(size_t)r * out_stride + (step_ptr ? *step_ptr : 0)] = (int64_t)idx;
}
cudaError_t launch_argmax_finalize(unsigned long long* best, int rows, int32_t* next_ids, int64_t* out64,
                                   int out_stride, const int* step_ptr, cudaStream_t s) {
  return launch_k(argmax_finalize_kernel, dim3((rows + 127) / 128), dim3(128), 0, s, best, rows, next_ids, out64, out_stride,
                  step_ptr);
}

// -------------------------------------------------------------------------- device scalars
__global__ void set_int_kernel(int* p, int v) { pdl_launch(); pdl_wait(); *p = v; }
__global__ void add_int_kernel(int* p, int v) { pdl_launch(); pdl_wait(); *p += v; }
cudaError_t launch_set_int(int* p, int v, cudaStream_t s) { return launch_k(set_int_kernel, dim3(1), dim3(1), 0, s, p, v); }
cudaError_t launch_add_int(int* p, int v, cudaStream_t s) { return launch_k(add_int_kernel, dim3(1), dim3(1), 0, s, p, v); }

// -------------------------------------------------------------------------- per-op RoPE / SwiGLU
// x [B, L, heads, HD] -> out, interleaved-pair rotation (llama3.py:41-76)
__global__ void rope_only_kernel(const float* __restrict__ x, const float* __restrict__ cos_tab,
                                 const float* __restrict__ sin_tab, int B, int L, int heads, int HD,
                                 const int* __restrict__ pos_ptr, float* __restrict__ out) {
  const int64_t npairs = (int64_t)B * L * heads * (HD / 2);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npairs) return;
  const int j = (int)(i % (HD / 2));
  const int64_t tok = i / ((int64_t)heads * (HD / 2));
  const int t = (int)(tok % L);
  const int pos = *pos_ptr + t;
  const float c = cos_tab[(size_t)pos * (HD / 2) + j], s = sin_tab[(size_t)pos * (HD / 2) + j];
  const float a = x[2 * i], b = x[2 * i + 1];
  out[2 * i] = a * c - b * s;
  out[2 * i + 1] = a * s + b * c;
}
cudaError_t launch_rope_only(const float* x, const float* cos_tab, const float* sin_tab, int B, int L, int heads,
                             int HD, const int* pos_ptr, float* out, cudaStream_t s) {
  const int64_t npairs = (int64_t)B * L * heads * (HD / 2);
  rope_only_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(x, cos_tab, sin_tab, B, L, heads, HD, pos_ptr, out);
  return cudaGetLastError();
}

__global__ void swiglu_kernel(const float* __restrict__ gate, const float* __restrict__ up, int64_t n,
                              float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = silu_ref(gate[i]) * up[i];
}
cudaError_t launch_swiglu(const float* gate, const float* up, int64_t n, float* out, cudaStream_t s) {
  swiglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(gate, up, n, out);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------- weight packing
// Copy a [rows, cols] fp32 matrix into rows dst_row0 + r * dst_row_stride of a packed matrix
// with leading dimension dst_ld (fused QKV: stride 1; interleaved gate/up: stride 2).
template <typename WT>
__global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int cols, WT* __restrict__ dst,
                                 int dst_row0, int dst_row_stride, int dst_ld) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[((size_t)dst_row0 + (size_t)r * dst_row_stride) * dst_ld + c] = from_f32<WT>(src[i]);
  }
}
cudaError_t launch_pack_rows(const float* src, int rows, int cols, void* dst, bool dst_bf16, int dst_row0,
                             int dst_row_stride, int dst_ld, cudaStream_t s) {
  const int64_t n = (int64_t)rows * cols;
  int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  if (dst_bf16) pack_rows_kernel<bf16><<<grid, 256, 0, s>>>(src, rows, cols, (bf16*)dst, dst_row0, dst_row_stride, dst_ld);
  else pack_rows_kernel<float><<<grid, 256, 0, s>>>(src, rows, cols, (float*)dst, dst_row0, dst_row_stride, dst_ld);
  return cudaGetLastError();
}

// Counter-based normal generator keyed on the GLOBAL (row, col) of the logical tensor, so that
// every tensor-parallel rank fills its slice of the same matrix a single GPU would build.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename WT>
__global__ void fill_random_kernel(WT* __restrict__ dst, int64_t rows, int64_t cols, int64_t ld_global,
                                   int64_t row0_global, int64_t col0_global, int dst_row0, int dst_row_stride,
                                   int dst_ld, uint64_t seed, uint32_t tensor_id, float scale, float bias) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    const uint64_t gidx = (uint64_t)(row0_global + r) * (uint64_t)ld_global + (uint64_t)(col0_global + c);
    const uint64_t h = splitmix64(splitmix64(seed ^ ((uint64_t)tensor_id << 40)) + gidx);
    const float u1 = ((uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777217.0f);  // (0, 1)
    const float u2 = (uint32_t)((h >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
    const float z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    dst[((size_t)dst_row0 + (size_t)r * dst_row_stride) * dst_ld + c] = from_f32<WT>(bias + scale * z);
  }
}
cudaError_t launch_fill_random(void* dst, bool dst_bf16, int64_t rows, int64_t cols, int64_t ld_global,
                               int64_t row0_global, int64_t col0_global, int dst_row0, int dst_row_stride,
                               int dst_ld, uint64_t seed, uint32_t tensor_id, float scale, float bias,
                               cudaStream_t s) {
  const int64_t n = rows * cols;
  int grid = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (grid < 1) grid = 1;
  if (dst_bf16)
    fill_random_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)dst, rows, cols, ld_global, row0_global, col0_global, dst_row0,
                                                  dst_row_stride, dst_ld, seed, tensor_id, scale, bias);
  else
    fill_random_kernel<float><<<grid, 256, 0, s>>>((float*)dst, rows, cols, ld_global, row0_global, col0_global,
                                                   dst_row0, dst_row_stride, dst_ld, seed, tensor_id, scale, bias);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------- cache layout conversion
// device layout [maxB, KVHN, M, HD]  <->  reference layout [B, T|M, KVHN, HD] (llama3.py:138-153)
template <typename KVT>
__global__ void cache_to_ref_kernel(const KVT* __restrict__ cache, int maxB, int KVHN, int M, int HD,
                                    float* __restrict__ out) {
  const int64_t n = (int64_t)maxB * KVHN * M * HD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % HD);
    const int h = (int)((i / HD) % KVHN);
    const int t = (int)((i / ((int64_t)HD * KVHN)) % M);
    const int b = (int)(i / ((int64_t)HD * KVHN * M));
    out[i] = to_f32(cache[(((size_t)b * KVHN + h) * M + t) * HD + d]);
  }
}
cudaError_t launch_cache_to_ref_layout(const void* cache, bool kv_bf16, int maxB, int KVHN, int M, int HD,
                                       float* out, cudaStream_t s) {
  if (kv_bf16) cache_to_ref_kernel<bf16><<<148 * 4, 256, 0, s>>>((const bf16*)cache, maxB, KVHN, M, HD, out);
  else cache_to_ref_kernel<float><<<148 * 4, 256, 0, s>>>((const float*)cache, maxB, KVHN, M, HD, out);
  return cudaGetLastError();
}

template <typename KVT>
__global__ void cache_from_ref_kernel(const float* __restrict__ in, int B, int T, int KVHN, int M, int HD,
                                      KVT* __restrict__ cache) {
  const int64_t n = (int64_t)B * T * KVHN * HD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % HD);
    const int h = (int)((i / HD) % KVHN);
    const int t = (int)((i / ((int64_t)HD * KVHN)) % T);
    const int b = (int)(i / ((int64_t)HD * KVHN * T));
    cache[(((size_t)b * KVHN + h) * M + t) * HD + d] = from_f32<KVT>(in[i]);
  }
}
cudaError_t launch_cache_from_ref_layout(const float* in, bool kv_bf16, int B, int T, int KVHN, int M, int HD,
                                         void* cache, cudaStream_t s) {
  if (kv_bf16) cache_from_ref_kernel<bf16><<<148 * 4, 256, 0, s>>>(in, B, T, KVHN, M, HD, (bf16*)cache);
  else cache_from_ref_kernel<float><<<148 * 4, 256, 0, s>>>(in, B, T, KVHN, M, HD, (float*)cache);
  return cudaGetLastError();
}

// bf16 -> fp32 (per-op test plumbing)
__global__ void unpack_bf16_kernel(const bf16* __restrict__ in, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}
cudaError_t launch_unpack_bf16(const bf16* in, int64_t n, float* out, cudaStream_t s) {
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  unpack_bf16_kernel<<<grid, 256, 0, s>>>(in, n, out);
  return cudaGetLastError();
}

// x += delta (bf16), 8 elements per thread: the residual add after a bf16 all-reduce of the partial
// projections (tensor-parallel prefill, llama3.py:253, 259)
__global__ void add_bf16_kernel(float* __restrict__ x, const bf16* __restrict__ d, int64_t n8) {
  pdl_launch();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = *reinterpret_cast<const uint4*>(d + i * 8);
    float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    a.x += __uint_as_float(r.x << 16); a.y += __uint_as_float(r.x & 0xffff0000u);
    a.z += __uint_as_float(r.y << 16); a.w += __uint_as_float(r.y & 0xffff0000u);
    b.x += __uint_as_float(r.z << 16); b.y += __uint_as_float(r.z & 0xffff0000u);
    b.z += __uint_as_float(r.w << 16); b.w += __uint_as_float(r.w & 0xffff0000u);
    *reinterpret_cast<float4*>(x + i * 8) = a;
    *reinterpret_cast<float4*>(x + i * 8 + 4) = b;
  }
}
cudaError_t launch_add_bf16(float* x, const bf16* d, int64_t n, cudaStream_t s) {
  const int64_t n8 = n / 8;
  const int grid = (int)std::min<int64_t>((n8 + 255) / 256, 148 * 8);
  return launch_k(add_bf16_kernel, dim3(grid), dim3(256), 0, s, x, d, n8);
}

// -------------------------------------------------------------------------- ragged batches
// Per-sequence bookkeeping of l3_generate_ragged.  step < 0: after the prefill (positions of the
// first decode step are set by the first advance); lastrow[b] = row of prompt b's last real token.
__global__ void ragged_setup_kernel(const int* __restrict__ len, int B, int Lmax, int32_t* __restrict__ lastrow,
                                    int* __restrict__ done) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch();
  pdl_wait();
  if (b >= B) return;
  lastrow[b] = b * Lmax + len[b] - 1;
  done[b] = 0;
}
// decode step i = ++scal[1]: sequence b runs at pos = len[b] + off + i (off 0: llama3.py:316-318, -1: llama3_simple.py:279)
__global__ void ragged_advance_kernel(int* __restrict__ scal, const int* __restrict__ len, int off, int B, int* __restrict__ rowpos) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch();
  pdl_wait();
  const int step = scal[1] + 1;
  if (b < B) rowpos[b] = len[b] + off + step;
  __syncthreads();
  if (b == 0) scal[1] = step;  // single block: every thread has read the old value
}
// per-sequence EOS: a finished sequence keeps emitting eos; the id that finishes it is kept
__global__ void ragged_eos_kernel(int32_t* __restrict__ next_ids, int* __restrict__ done, int eos, int B, int64_t* __restrict__ tokens,
                                  int stride, const int* __restrict__ step_ptr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch();
  pdl_wait();
  if (b >= B) return;
  if (done[b]) {
    next_ids[b] = eos;
    tokens[(size_t)b * stride + *step_ptr] = eos;
  } else if (next_ids[b] == eos) {
    done[b] = 1;
  }
}
cudaError_t launch_ragged_setup(const int* len, int B, int Lmax, int32_t* lastrow, int* done, cudaStream_t s) {
  return launch_k(ragged_setup_kernel, dim3((B + 255) / 256), dim3(256), 0, s, len, B, Lmax, lastrow, done);
}
cudaError_t launch_ragged_advance(int* scal, const int* len, int off, int B, int* rowpos, cudaStream_t s) {
  return launch_k(ragged_advance_kernel, dim3(1), dim3(1024), 0, s, scal, len, off, B, rowpos);  // B <= 1024
}
cudaError_t launch_ragged_eos(int32_t* next_ids, int* done, int eos, int B, int64_t* tokens, int stride, const int* step_ptr,
                              cudaStream_t s) {
  return launch_k(ragged_eos_kernel, dim3((B + 255) / 256), dim3(256), 0, s, next_ids, done, eos, B, tokens, stride, step_ptr);
}
