// Design-input micro-benchmarks for the cluster-resident decode kernel.  Not on any product path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/ubench scripts/ubench/ubench.cu
// One JSON line per measurement:
//   clusters   - cudaOccupancyMaxActiveClusters per cluster size at a given dynamic shared-memory size, and
//                whether a cluster of that size launches and exchanges data through DSMEM correctly
//   mma_sync   - legacy mma.sync m16n8k8 TF32 / m16n8k16 BF16 cycles per instruction per SM by warps per SM
//   ffma       - FFMA issue rate (sanity)
//   ingress    - cp.async.bulk global -> shared bandwidth per SM and chip-wide, L2-resident and HBM streaming
//   dsmem      - st.shared::cluster bandwidth and barrier.cluster round-trip time
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e__ = (x);                                                                     \
    if (e__ != cudaSuccess) {                                                                  \
      printf("{\"error\": \"%s at line %d: %s\"}\n", #x, __LINE__, cudaGetErrorString(e__)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------ clusters
// Every CTA writes (rank + 1) * 1000 + j into slot [rank][j] of EVERY CTA of its cluster, barrier, then each CTA
// checks all slots and records smid / start / end times.
__global__ void cluster_probe_kernel(int* ok, unsigned* sm_of, unsigned long long* t0s, unsigned long long* t1s, int spin_us) {
  extern __shared__ float4 sm4[];
  float* slots = reinterpret_cast<float*>(sm4);  // [16][64]
  const uint32_t rank = cluster_ctarank(), n = cluster_nctarank();
  const unsigned long long t0 = gtime();
  cluster_sync();  // every CTA of the cluster is running: its shared memory may be written
  const uint32_t base = smem_u32(slots);
  for (uint32_t p = 0; p < n; ++p) {
    const uint32_t dst = mapa(base, p) + (rank * 64 + threadIdx.x * 4) * 4;
    if (threadIdx.x < 16) st_cluster_f4(dst, make_float4(rank * 1000.f + threadIdx.x * 4, rank * 1000.f + threadIdx.x * 4 + 1,
                                                         rank * 1000.f + threadIdx.x * 4 + 2, rank * 1000.f + threadIdx.x * 4 + 3));
  }
  cluster_sync();
  int good = 1;
  for (uint32_t p = 0; p < n; ++p)
    for (int j = threadIdx.x; j < 64; j += blockDim.x)
      if (slots[p * 64 + j] != p * 1000.f + j) good = 0;
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) {}
  cluster_sync();  // nobody exits while a peer may still read its shared memory
  if (!good) atomicExch(ok, 0);
  if (threadIdx.x == 0) { sm_of[blockIdx.x] = smid(); t0s[blockIdx.x] = t0; t1s[blockIdx.x] = gtime(); }
}

static int run_clusters(int smem_bytes) {
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaFuncSetAttribute(cluster_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int *d_ok; unsigned* d_sm; unsigned long long *d_t0, *d_t1;
  CK(cudaMalloc(&d_ok, 4)); CK(cudaMalloc(&d_sm, 4096 * 4)); CK(cudaMalloc(&d_t0, 4096 * 8)); CK(cudaMalloc(&d_t1, 4096 * 8));
  for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 12, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(cs);
    int maxc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, cluster_probe_kernel, &cfg);
    if (e != cudaSuccess) { printf("{\"bench\": \"clusters\", \"size\": %d, \"smem\": %d, \"occupancy_error\": \"%s\"}\n", cs, smem_bytes, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    // launch maxc + 2 clusters that each spin 200 us: concurrency = clusters whose start precedes the first end
    const int ncl = maxc > 0 ? maxc + 2 : 1;
    cfg.gridDim = dim3(ncl * cs);
    int one = 1;
    CK(cudaMemcpy(d_ok, &one, 4, cudaMemcpyHostToDevice));
    e = cudaLaunchKernelEx(&cfg, cluster_probe_kernel, d_ok, d_sm, d_t0, d_t1, 200);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"bench\": \"clusters\", \"size\": %d, \"smem\": %d, \"max_active\": %d, \"launch_error\": \"%s\"}\n", cs, smem_bytes, maxc, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    int ok = 0;
    std::vector<unsigned long long> t0(ncl * cs), t1(ncl * cs);
    std::vector<unsigned> smv(ncl * cs);
    CK(cudaMemcpy(&ok, d_ok, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t0.data(), d_t0, t0.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t1.data(), d_t1, t1.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(smv.data(), d_sm, smv.size() * 4, cudaMemcpyDeviceToHost));
    unsigned long long first_end = ~0ull;
    for (auto v : t1) first_end = v < first_end ? v : first_end;
    int conc = 0;
    for (int c = 0; c < ncl; ++c) conc += t0[c * cs] < first_end;
    std::vector<int> used(256, 0);
    int distinct = 0;
    for (size_t i = 0; i < smv.size(); ++i) if (t0[i] < first_end && !used[smv[i] & 255]++) distinct++;
    printf("{\"bench\": \"clusters\", \"size\": %d, \"smem\": %d, \"sms\": %d, \"max_active_clusters\": %d, \"launched\": %d, \"concurrent\": %d, "
           "\"distinct_sms_first_wave\": %d, \"dsmem_ok\": %d}\n", cs, smem_bytes, nsm, maxc, ncl, conc, distinct, ok);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ mma.sync / FFMA
template <int NACC>
__global__ void mma_tf32_kernel(int iters, float* sink, long long* cyc) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NACC>
__global__ void mma_bf16_kernel(int iters, float* sink, long long* cyc) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NACC>
__global__ void ffma_kernel(int iters, float* sink, long long* cyc) {
  float c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = (float)i;
  float a = threadIdx.x * 1e-3f, b = 1.0001f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fmaf(c[i], b, a);
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 12345.678f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int run_math() {
  float* sink; long long* d_cyc;
  CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&d_cyc, 1024 * 8));
  const int iters = 4096;
  constexpr int NACC = 8;
  for (int ctas : {1, 148}) {
    for (int warps : {4, 8, 16}) {
      for (int which = 0; which < 3; ++which) {
        if (which == 0) mma_tf32_kernel<NACC><<<ctas, warps * 32>>>(iters, sink, d_cyc);
        else if (which == 1) mma_bf16_kernel<NACC><<<ctas, warps * 32>>>(iters, sink, d_cyc);
        else ffma_kernel<NACC><<<ctas, warps * 32>>>(iters, sink, d_cyc);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(ctas);
        CK(cudaMemcpy(h.data(), d_cyc, ctas * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        const double instr_per_sm = (double)iters * NACC * warps;
        const double cyc_per_instr = (double)mx / instr_per_sm;
        const double mac = which == 0 ? 16 * 8 * 8 : which == 1 ? 16 * 8 * 16 : 32;
        printf("{\"bench\": \"%s\", \"ctas\": %d, \"warps_per_sm\": %d, \"sm_cycles_per_warp_instr\": %.3f, \"mac_per_clk_per_sm\": %.1f}\n",
               which == 0 ? "mma_sync_tf32_m16n8k8" : which == 1 ? "mma_sync_bf16_m16n8k16" : "ffma", ctas, warps, cyc_per_instr,
               mac / cyc_per_instr);
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ bulk ingress
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// One thread per CTA keeps `stages` bulk copies of `chunk` bytes in flight; CTA c reads region [c * per_cta, +per_cta)
// (per_cta = 0: every CTA reads the same `span` bytes - the L2-resident case).
__global__ void ingress_kernel(const uint8_t* src, size_t per_cta, size_t span, int chunk, int stages, int nchunks, long long* cyc,
                               unsigned long long* ns) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ __align__(8) unsigned long long bars[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + (per_cta ? (size_t)blockIdx.x * per_cta : 0);
    const size_t lim = per_cta ? per_cta : span;
    const unsigned long long g0 = gtime();
    const long long t0 = clock64();
    for (int i = 0; i < nchunks + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(smem_u32(&bars[s]), ((i / stages) - 1) & 1);
      if (i < nchunks) {
        // CTAs start at different offsets of a shared region so that they do not all hit the same lines at once
        const size_t off = (((size_t)i + (per_cta ? 0 : (size_t)blockIdx.x * 7)) * (size_t)chunk) % (lim - chunk + 1);
        mbar_expect_tx(smem_u32(&bars[s]), chunk);
        bulk_g2s(smem_u32(ring + (size_t)s * chunk), base + (off & ~(size_t)15), chunk, smem_u32(&bars[s]));
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
    ns[blockIdx.x] = gtime() - g0;
  }
}

static int run_ingress() {
  const size_t total = (size_t)4 << 30;  // 4 GB source: HBM streaming regions
  uint8_t* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  long long* d_cyc; unsigned long long* d_ns;
  CK(cudaMalloc(&d_cyc, 1024 * 8)); CK(cudaMalloc(&d_ns, 1024 * 8));
  CK(cudaFuncSetAttribute(ingress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int ctas : {1, 16, 148}) {
    for (int mode = 0; mode < 2; ++mode) {       // 0 = L2-resident shared 16 MB region, 1 = HBM streaming
      for (int chunk : {4096, 16384, 32768, 65536}) {
        for (int stages : {2, 3, 4, 8}) {
          if ((size_t)chunk * stages > 196 * 1024) continue;
          const size_t per_cta = mode ? ((total / ctas) & ~(size_t)4095) : 0;
          const size_t span = (size_t)16 << 20;
          const size_t bytes_per_cta = mode ? ((size_t)24 << 20) : ((size_t)32 << 20);
          const int nchunks = (int)(bytes_per_cta / chunk);
          if (mode && (size_t)nchunks * chunk > per_cta) continue;
          for (int rep = 0; rep < 2; ++rep) {  // first repetition warms L2 in mode 0
            ingress_kernel<<<ctas, 32, (size_t)chunk * stages>>>(src, per_cta, span, chunk, stages, nchunks, d_cyc, d_ns);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
          }
          std::vector<unsigned long long> ns(ctas);
          CK(cudaMemcpy(ns.data(), d_ns, ctas * 8, cudaMemcpyDeviceToHost));
          unsigned long long mx = 0;
          for (auto v : ns) mx = v > mx ? v : mx;
          const double gbs_sm = (double)nchunks * chunk / (double)mx;  // bytes / ns = GB/s
          printf("{\"bench\": \"ingress\", \"source\": \"%s\", \"ctas\": %d, \"chunk\": %d, \"stages\": %d, \"GBps_per_sm\": %.1f, \"GBps_total\": %.1f}\n",
                 mode ? "hbm" : "l2", ctas, chunk, stages, gbs_sm, gbs_sm * ctas);
        }
      }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ DSMEM
// Every CTA of a cluster pushes `bytes` to each peer with st.shared::cluster.v4 (all threads), then barrier; repeated.
__global__ void dsmem_kernel(int bytes, int reps, long long* cyc_push, long long* cyc_bar) {
  extern __shared__ float4 buf4[];
  const uint32_t rank = cluster_ctarank(), n = cluster_nctarank();
  const uint32_t base = smem_u32(buf4);
  cluster_sync();
  long long push = 0, bar = 0;
  for (int r = 0; r < reps; ++r) {
    const long long t0 = clock64();
    for (uint32_t p = 1; p < n; ++p) {
      const uint32_t dst = mapa(base, (rank + p) % n) + rank * bytes;
      for (int o = threadIdx.x * 16; o < bytes; o += blockDim.x * 16) st_cluster_f4(dst + o, make_float4(1.f, 2.f, 3.f, (float)r));
    }
    const long long t1 = clock64();
    cluster_sync();
    const long long t2 = clock64();
    push += t1 - t0;
    bar += t2 - t1;
  }
  if (threadIdx.x == 0) { cyc_push[blockIdx.x] = push / reps; cyc_bar[blockIdx.x] = bar / reps; }
}
__global__ void cluster_bar_kernel(int reps, long long* cyc) {
  cluster_sync();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) cluster_sync();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (clock64() - t0) / reps;
}

static int run_dsmem() {
  long long *d_a, *d_b;
  CK(cudaMalloc(&d_a, 1024 * 8)); CK(cudaMalloc(&d_b, 1024 * 8));
  CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  for (int cs : {2, 4, 6, 8}) {
    for (int threads : {256, 512}) {
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(threads);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cfg.gridDim = dim3(cs * 8);
      {
        cfg.dynamicSmemBytes = 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_bar_kernel, 64, d_a);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"bench\": \"cluster_barrier\", \"size\": %d, \"error\": \"%s\"}\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        long long h[64];
        CK(cudaMemcpy(h, d_a, cs * 8 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < cs * 8; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("{\"bench\": \"cluster_barrier\", \"size\": %d, \"threads\": %d, \"cycles\": %lld}\n", cs, threads, mx);
      }
      for (int bytes : {2304, 8192}) {
        cfg.dynamicSmemBytes = 8 * bytes;
        cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, bytes, 32, d_a, d_b);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"bench\": \"dsmem_push\", \"size\": %d, \"error\": \"%s\"}\n", cs, cudaGetErrorString(e)); cudaGetLastError(); continue; }
        long long ha[64], hb[64];
        CK(cudaMemcpy(ha, d_a, cs * 8 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hb, d_b, cs * 8 * 8, cudaMemcpyDeviceToHost));
        long long pa = 0, pb = 0;
        for (int i = 0; i < cs * 8; ++i) { pa = ha[i] > pa ? ha[i] : pa; pb = hb[i] > pb ? hb[i] : pb; }
        printf("{\"bench\": \"dsmem_push\", \"size\": %d, \"threads\": %d, \"bytes_to_each_peer\": %d, \"issue_cycles\": %lld, \"barrier_after_cycles\": %lld, "
               "\"bytes_per_clk_out\": %.1f}\n", cs, threads, bytes, pa, pb, (double)bytes * (cs - 1) / (double)(pa + pb));
      }
    }
  }
  return 0;
}

// Variant: NT issuing threads (one per warp), each with its own ring of `stages` x `chunk` bytes.
__global__ void ingress_multi_kernel(const uint8_t* src, size_t per_cta, int chunk, int stages, int nchunks, unsigned long long* ns) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ __align__(8) unsigned long long bars[64];
  const int nt = blockDim.x / 32, w = threadIdx.x / 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages * nt; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long g0 = gtime();
  if ((threadIdx.x & 31) == 0) {
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta + (size_t)w * (per_cta / nt & ~(size_t)4095);
    for (int i = 0; i < nchunks + stages; ++i) {
      const int s = i % stages;
      const uint32_t bar = smem_u32(&bars[w * stages + s]);
      if (i >= stages) mbar_wait(bar, ((i / stages) - 1) & 1);
      if (i < nchunks) {
        mbar_expect_tx(bar, chunk);
        bulk_g2s(smem_u32(ring + (size_t)(w * stages + s) * chunk), base + (size_t)i * chunk, chunk, bar);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) ns[blockIdx.x] = gtime() - g0;
}
// Variant: plain 16-byte loads by every thread (8 in flight per thread), stored to shared memory.
__global__ void ingress_ldg_kernel(const uint8_t* src, size_t per_cta, size_t bytes, unsigned long long* ns) {
  extern __shared__ __align__(128) uint8_t ring[];
  const uint4* p = reinterpret_cast<const uint4*>(src + (size_t)blockIdx.x * per_cta);
  uint4* sm = reinterpret_cast<uint4*>(ring);
  const size_t n16 = bytes / 16;
  const unsigned long long g0 = gtime();
  for (size_t i = threadIdx.x; i + 7 * blockDim.x < n16; i += 8 * blockDim.x) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w) : "l"(p + i + j * blockDim.x));
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[(threadIdx.x + j * blockDim.x) & 4095] = v[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) ns[blockIdx.x] = gtime() - g0;
}
// Variant: NL issuing LANES of ONE warp, each with its own ring (does the per-issuer limit apply per thread or per warp?)
__global__ void ingress_lanes_kernel(const uint8_t* src, size_t per_cta, int chunk, int stages, int nchunks, int nl, unsigned long long* ns) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ __align__(8) unsigned long long bars[64];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages * nl; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned long long g0 = gtime();
  const int w = threadIdx.x;
  if (w < nl) {
    const uint8_t* base = src + (size_t)blockIdx.x * per_cta + (size_t)w * (per_cta / nl & ~(size_t)4095);
    for (int i = 0; i < nchunks + stages; ++i) {
      const int s = i % stages;
      const uint32_t bar = smem_u32(&bars[w * stages + s]);
      if (i >= stages) mbar_wait(bar, ((i / stages) - 1) & 1);
      if (i < nchunks) {
        mbar_expect_tx(bar, chunk);
        bulk_g2s(smem_u32(ring + (size_t)(w * stages + s) * chunk), base + (size_t)i * chunk, chunk, bar);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) ns[blockIdx.x] = gtime() - g0;
}
static int run_ingress3() {
  const size_t total = (size_t)4 << 30;
  uint8_t* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  unsigned long long* d_ns;
  CK(cudaMalloc(&d_ns, 1024 * 8));
  CK(cudaFuncSetAttribute(ingress_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int ctas : {1, 148}) {
    const size_t per_cta = (total / ctas) & ~(size_t)65535;
    const size_t bytes_per_cta = (size_t)24 << 20;
    for (int nl : {1, 2, 4, 8}) {
      for (int chunk : {8192, 16384}) {
        const int stages = 2;
        if ((size_t)chunk * stages * nl > 192 * 1024) continue;
        const int nchunks = (int)(bytes_per_cta / nl / chunk);
        ingress_lanes_kernel<<<ctas, 32, (size_t)chunk * stages * nl>>>(src, per_cta, chunk, stages, nchunks, nl, d_ns);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> ns(ctas);
        CK(cudaMemcpy(ns.data(), d_ns, ctas * 8, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (auto v : ns) mx = v > mx ? v : mx;
        const double gbs = (double)nchunks * chunk * nl / (double)mx;
        printf("{\"bench\": \"ingress_lanes\", \"ctas\": %d, \"issuing_lanes_of_one_warp\": %d, \"chunk\": %d, \"stages_each\": %d, \"GBps_per_sm\": %.1f, \"GBps_total\": %.1f}\n",
               ctas, nl, chunk, stages, gbs, gbs * ctas);
      }
    }
  }
  return 0;
}

static int run_ingress2() {
  const size_t total = (size_t)4 << 30;
  uint8_t* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  unsigned long long* d_ns;
  CK(cudaMalloc(&d_ns, 1024 * 8));
  CK(cudaFuncSetAttribute(ingress_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(ingress_ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int ctas : {1, 132, 148}) {
    const size_t per_cta = (total / ctas) & ~(size_t)65535;
    const size_t bytes_per_cta = (size_t)24 << 20;
    for (int nt : {2, 4, 8}) {
      for (int chunk : {4096, 8192, 16384}) {
        for (int stages : {1, 2, 3}) {
          if ((size_t)chunk * stages * nt > 192 * 1024) continue;
          const int nchunks = (int)(bytes_per_cta / nt / chunk);
          ingress_multi_kernel<<<ctas, nt * 32, (size_t)chunk * stages * nt>>>(src, per_cta, chunk, stages, nchunks, d_ns);
          CK(cudaGetLastError());
          CK(cudaDeviceSynchronize());
          std::vector<unsigned long long> ns(ctas);
          CK(cudaMemcpy(ns.data(), d_ns, ctas * 8, cudaMemcpyDeviceToHost));
          unsigned long long mx = 0;
          for (auto v : ns) mx = v > mx ? v : mx;
          const double gbs = (double)nchunks * chunk * nt / (double)mx;
          printf("{\"bench\": \"ingress_multi\", \"ctas\": %d, \"issuers\": %d, \"chunk\": %d, \"stages_each\": %d, \"GBps_per_sm\": %.1f, \"GBps_total\": %.1f}\n",
                 ctas, nt, chunk, stages, gbs, gbs * ctas);
        }
      }
    }
    for (int threads : {256, 512, 1024}) {
      ingress_ldg_kernel<<<ctas, threads, 64 * 1024>>>(src, per_cta, bytes_per_cta, d_ns);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      std::vector<unsigned long long> ns(ctas);
      CK(cudaMemcpy(ns.data(), d_ns, ctas * 8, cudaMemcpyDeviceToHost));
      unsigned long long mx = 0;
      for (auto v : ns) mx = v > mx ? v : mx;
      const double gbs = (double)bytes_per_cta / (double)mx;
      printf("{\"bench\": \"ingress_ldg\", \"ctas\": %d, \"threads\": %d, \"GBps_per_sm\": %.1f, \"GBps_total\": %.1f}\n", ctas, threads, gbs, gbs * ctas);
    }
  }
  return 0;
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  const char* what = argc > 1 ? argv[1] : "all";
  auto want = [&](const char* s) { return !strcmp(what, "all") || !strcmp(what, s); };
  int rc = 0;
  if (want("clusters")) { rc |= run_clusters(64 * 1024); rc |= run_clusters(200 * 1024); }
  if (want("math")) rc |= run_math();
  if (want("dsmem")) rc |= run_dsmem();
  if (want("ingress")) rc |= run_ingress();
  if (want("ingress2")) rc |= run_ingress2();
  if (want("ingress3")) rc |= run_ingress3();
  return rc;
}
