// Per-op C-ABI entry points (include/llama3_b200.h, "per-op entry points"): each runs the
// very kernels the model path launches, on caller-supplied host buffers, so that parity
// tests can check one reference function at a time.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "../../include/llama3_b200.h"
#include "common.cuh"
#include "gemm_tc.h"

namespace {
struct Scratch {  // frees everything it handed out
  std::vector<void*> ptrs;
  cudaStream_t s = nullptr;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
    if (s) cudaStreamDestroy(s);
  }
  template <typename T> T* dev(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return (T*)p;
  }
  template <typename T> T* up(const T* host, size_t n) {
    T* p = dev<T>(n);
    if (p) cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    return p;
  }
};
int finish(Scratch& sc, cudaError_t launch_err) {
  if (launch_err != cudaSuccess) { fprintf(stderr, "l3_op launch: %s\n", cudaGetErrorString(launch_err)); return L3_ECUDA; }
  cudaError_t e = cudaStreamSynchronize(sc.s);
  if (e != cudaSuccess) { fprintf(stderr, "l3_op sync: %s\n", cudaGetErrorString(e)); return L3_ECUDA; }
  return L3_OK;
}
int begin(Scratch& sc, int device) {
  if (cudaSetDevice(device) != cudaSuccess) return L3_ECUDA;
  if (cudaStreamCreate(&sc.s) != cudaSuccess) return L3_ECUDA;
  return L3_OK;
}
}  // namespace

extern "C" int l3_op_rmsnorm(int device, const float* x, const float* w, float eps, int rows, int dim, float* out) {
  if (dim % 4) return L3_EINVAL;
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  float* dx = sc.up(x, (size_t)rows * dim);
  float* dw = sc.up(w, dim);
  float* dout = sc.dev<float>((size_t)rows * dim);
  if (!dx || !dw || !dout) return L3_ENOMEM;
  int rc = finish(sc, launch_rmsnorm(dx, dw, eps, rows, dim, 1, 0, dout, nullptr, nullptr, sc.s));
  if (rc == L3_OK) cudaMemcpy(out, dout, (size_t)rows * dim * 4, cudaMemcpyDeviceToHost);
  return rc;
}

extern "C" int l3_op_linear(int device, const float* x, const float* w, int rows, int n, int k, int path, int w_bf16,
                            float* out) {
  if (k % 8) return L3_EINVAL;
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  float* dx = sc.up(x, (size_t)rows * k);
  float* dw32 = sc.up(w, (size_t)n * k);
  float* dout = sc.dev<float>((size_t)rows * n);
  if (!dx || !dw32 || !dout) return L3_ENOMEM;
  void* dw = dw32;
  if (w_bf16) {
    bf16* dwb = sc.dev<bf16>((size_t)n * k);
    if (!dwb) return L3_ENOMEM;
    cudaError_t e = launch_pack_rows(dw32, n, k, dwb, true, 0, 1, k, sc.s);
    if (e != cudaSuccess) return L3_ECUDA;
    dw = dwb;
  }
  LinearArgs a{};
  a.W = dw; a.x = dx; a.rows = rows; a.N = n; a.K = k; a.src_mul = 1; a.src_add = 0;
  a.epi = EPI_STORE; a.e.out = dout; a.e.ld_out = n;
  cudaError_t e;
  if (path == 0) path = linear_rows_supported(rows, k) ? 1 : 2;
  if (path == 1) {
    if (!linear_rows_supported(rows, k)) return L3_EINVAL;
    e = launch_linear_rows(a, w_bf16 != 0, false, sc.s);
  } else if (path == 2) {
    e = launch_linear_simt(a, w_bf16 != 0, false, sc.s);
  } else if (path == 3 || path == 4) {  // tcgen05: bf16 operands, or the 3xTF32 split of fp32 operands (4: swapped roles)
    if (!tc_gemm_supported(k)) return L3_EINVAL;
    TcGemmArgs t{};
    t.rows = rows; t.N = n; t.K = k; t.epi = EPI_STORE; t.e = a.e; t.bn = 0;
    if (w_bf16) {
      bf16* dxb = sc.dev<bf16>((size_t)rows * k);
      if (!dxb) return L3_ENOMEM;
      if (launch_pack_rows(dx, rows, k, dxb, true, 0, 1, k, sc.s) != cudaSuccess) return L3_ECUDA;
      t.kind = TC_BF16; t.A[0] = dxb; t.W[0] = dw;
    } else {
      float* xh = sc.dev<float>((size_t)rows * k); float* xl = sc.dev<float>((size_t)rows * k);
      float* wh = sc.dev<float>((size_t)n * k); float* wl = sc.dev<float>((size_t)n * k);
      if (!xh || !xl || !wh || !wl) return L3_ENOMEM;
      if (launch_split_tf32(dx, xh, xl, (int64_t)rows * k, sc.s) != cudaSuccess) return L3_ECUDA;
      if (launch_split_tf32(dw32, wh, wl, (int64_t)n * k, sc.s) != cudaSuccess) return L3_ECUDA;
      t.kind = TC_TF32X3; t.A[0] = xh; t.A[1] = xl; t.W[0] = wh; t.W[1] = wl;
    }
    if (path == 4 && !gemm_swap_supported(rows, n)) return L3_EINVAL;
    // K-split scratch, as the model path owns it (l3_api.cu linear()): short-and-wide shapes split along K
    const size_t part_bytes = (size_t)32 << 20;
    t.part = sc.dev<float>(part_bytes / 4); t.part_bytes = part_bytes;
    t.tile_cnt = sc.dev<int>(1024); t.tile_cnt_len = 1024;
    if (!t.part || !t.tile_cnt) return L3_ENOMEM;
    if (cudaMemsetAsync(t.tile_cnt, 0, 1024 * sizeof(int), sc.s) != cudaSuccess) return L3_ECUDA;
    e = path == 4 ? launch_gemm_swap(t, sc.s) : launch_gemm_tc(t, sc.s);
    int rc = finish(sc, e);
    tc_forget_maps();  // the scratch buffers are about to be freed
    if (rc == L3_OK) cudaMemcpy(out, dout, (size_t)rows * n * 4, cudaMemcpyDeviceToHost);
    return rc;
  } else {
    return L3_EINVAL;
  }
  int rc = finish(sc, e);
  if (rc == L3_OK) cudaMemcpy(out, dout, (size_t)rows * n * 4, cudaMemcpyDeviceToHost);
  return rc;
}

extern "C" int l3_op_rope(int device, const float* x, const double* cos_tab, const double* sin_tab, int B, int L,
                          int heads, int head_dim, int start_pos, float* out) {
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  const size_t n = (size_t)B * L * heads * head_dim, nt = (size_t)(start_pos + L) * (head_dim / 2);
  std::vector<float> c(nt), s(nt);
  for (size_t i = 0; i < nt; ++i) { c[i] = (float)cos_tab[i]; s[i] = (float)sin_tab[i]; }
  float* dx = sc.up(x, n);
  float* dc = sc.up(c.data(), nt);
  float* ds = sc.up(s.data(), nt);
  float* dout = sc.dev<float>(n);
  int* dpos = sc.up(&start_pos, 1);
  if (!dx || !dc || !ds || !dout || !dpos) return L3_ENOMEM;
  int rc = finish(sc, launch_rope_only(dx, dc, ds, B, L, heads, head_dim, dpos, dout, sc.s));
  if (rc == L3_OK) cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost);
  return rc;
}

extern "C" int l3_op_swiglu(int device, const float* gate, const float* up, int64_t n, float* out) {
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  float* dg = sc.up(gate, (size_t)n);
  float* du = sc.up(up, (size_t)n);
  float* dout = sc.dev<float>((size_t)n);
  if (!dg || !du || !dout) return L3_ENOMEM;
  int rc = finish(sc, launch_swiglu(dg, du, n, dout, sc.s));
  if (rc == L3_OK) cudaMemcpy(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost);
  return rc;
}

extern "C" int l3_op_attention(int device, const float* q, const float* k, const float* v, int B, int L, int n_heads,
                               int n_kv_heads, int head_dim, int start_pos, int kv_bf16, int nsplit, float* out) {
  if (!attn_head_dim_supported(head_dim) || n_heads % n_kv_heads) return L3_EINVAL;
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  const int T = start_pos + L;
  const size_t nq = (size_t)B * L * n_heads * head_dim, nkv = (size_t)B * T * n_kv_heads * head_dim;
  float* dq = sc.up(q, nq);
  float* dk = sc.up(k, nkv);
  float* dv = sc.up(v, nkv);
  float* dout = sc.dev<float>(nq);
  int* dpos = sc.up(&start_pos, 1);
  const size_t kvb = kv_bf16 ? 2 : 4;
  void* ck = sc.dev<char>(nkv * kvb);
  void* cv = sc.dev<char>(nkv * kvb);
  if (nsplit < 1) nsplit = (L == 1 && T >= 64) ? 4 : 1;
  if (L > 1) nsplit = 1;
  float* po = sc.dev<float>((size_t)B * n_heads * nsplit * head_dim);
  float* pml = sc.dev<float>((size_t)B * n_heads * nsplit * 2);
  if (!dq || !dk || !dv || !dout || !dpos || !ck || !cv || !po || !pml) return L3_ENOMEM;
  cudaError_t e = launch_cache_from_ref_layout(dk, kv_bf16 != 0, B, T, n_kv_heads, T, head_dim, ck, sc.s);
  if (e == cudaSuccess) e = launch_cache_from_ref_layout(dv, kv_bf16 != 0, B, T, n_kv_heads, T, head_dim, cv, sc.s);
  if (e != cudaSuccess) return L3_ECUDA;
  AttnArgs a{};
  a.q = dq; a.cache_k = ck; a.cache_v = cv; a.out = dout; a.pos_ptr = dpos;
  a.B = B; a.L = L; a.HN = n_heads; a.KVHN = n_kv_heads; a.HD = head_dim; a.M = T;
  a.part_o = po; a.part_ml = pml; a.nsplit = nsplit;
  a.force_exact = kv_bf16 == 3;  // 3: bf16 cache through the exact lane-group kernels (1: the model's own choice)
  if (kv_bf16 == 2) {  // bf16 tensor-core flash prefill (attention_tc.cu): q and the output travel as bf16
    if (L == 1 || !attn_prefill_tc_supported(head_dim)) return L3_EINVAL;
    bf16* q16 = sc.dev<bf16>(nq);
    bf16* o16 = sc.dev<bf16>(nq);
    if (!q16 || !o16) return L3_ENOMEM;
    if (launch_pack_rows(dq, B * L, n_heads * head_dim, q16, true, 0, 1, n_heads * head_dim, sc.s) != cudaSuccess) return L3_ECUDA;
    a.out = nullptr; a.out_bf16 = o16; a.cache_rows = B * n_kv_heads * T;
    e = launch_attn_prefill_tc(a, q16, sc.s);
    if (e == cudaSuccess) e = launch_unpack_bf16(o16, (int64_t)nq, dout, sc.s);
    int rc2 = finish(sc, e);
    tc_forget_maps();
    if (rc2 == L3_OK) cudaMemcpy(out, dout, nq * 4, cudaMemcpyDeviceToHost);
    return rc2;
  }
  e = (L == 1) ? launch_attn_decode(a, kv_bf16 != 0, sc.s) : launch_attn_prefill(a, kv_bf16 != 0, sc.s);
  int rc = finish(sc, e);
  tc_forget_maps();  // the tensor-core decode kernel maps the scratch caches, which are about to be freed
  if (rc == L3_OK) cudaMemcpy(out, dout, nq * 4, cudaMemcpyDeviceToHost);
  return rc;
}

extern "C" int l3_op_argmax(int device, const float* logits, int rows, int n, int64_t* out) {
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  float* dl = sc.up(logits, (size_t)rows * n);
  int64_t* dout = sc.dev<int64_t>(rows);
  if (!dl || !dout) return L3_ENOMEM;
  int rc = finish(sc, launch_argmax(dl, rows, n, nullptr, dout, 1, nullptr, sc.s));
  if (rc == L3_OK) cudaMemcpy(out, dout, (size_t)rows * 8, cudaMemcpyDeviceToHost);
  return rc;
}


// Micro-benchmark of the row-streaming GEMV at one shape: rotates over enough weight copies to
// exceed the L2, `iters` launches back to back on one stream; returns the average ms per launch.
extern "C" int l3_bench_gemv(int device, int n, int k, int w_bf16, int rows, int iters, float* avg_ms) {
  if (!linear_rows_supported(rows, k) || iters < 1) return L3_EINVAL;
  Scratch sc;
  if (begin(sc, device)) return L3_ECUDA;
  const size_t es = w_bf16 ? 2 : 4, wbytes = (size_t)n * k * es;
  int nbuf = (int)(((size_t)320 << 20) / wbytes) + 1;
  if (nbuf < 2) nbuf = 2;
  if (nbuf > 64) nbuf = 64;
  std::vector<void*> W(nbuf);
  for (auto& w : W) {
    w = sc.dev<char>(wbytes);
    if (!w) return L3_ENOMEM;
    cudaMemsetAsync(w, 0, wbytes, sc.s);
  }
  float* dx = sc.dev<float>((size_t)rows * k);
  float* dout = sc.dev<float>((size_t)rows * n);
  float* dg = sc.dev<float>(k);
  if (!dx || !dout || !dg) return L3_ENOMEM;
  cudaMemsetAsync(dx, 0, (size_t)rows * k * 4, sc.s);
  cudaMemsetAsync(dg, 0, (size_t)k * 4, sc.s);
  LinearArgs a{};
  a.x = dx; a.rows = rows; a.N = n; a.K = k; a.src_mul = 1; a.src_add = 0; a.norm_w = dg; a.eps = 1e-6f;
  a.epi = EPI_STORE; a.e.out = dout; a.e.ld_out = n;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < nbuf && e == cudaSuccess; ++i) { a.W = W[i]; e = launch_linear_rows(a, w_bf16 != 0, false, sc.s); }
  cudaEventRecord(e0, sc.s);
  for (int i = 0; i < iters && e == cudaSuccess; ++i) { a.W = W[i % nbuf]; e = launch_linear_rows(a, w_bf16 != 0, false, sc.s); }
  cudaEventRecord(e1, sc.s);
  int rc = finish(sc, e);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *avg_ms = ms / iters;
  return rc;
}


// Debug: enable/disable the tcgen05 GEMM timeline stamps and fetch the last 64 clock64 values.
extern "C" int l3_debug_tc_timeline(int device, int enable, uint64_t* out64) {
  if (cudaSetDevice(device) != cudaSuccess) return L3_ECUDA;
  return tc_debug_timeline(enable, (unsigned long long*)out64) == 0 ? L3_OK : L3_ECUDA;
}
