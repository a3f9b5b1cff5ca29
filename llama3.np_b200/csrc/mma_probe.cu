// Micro-benchmark, not on any product path: what does one tcgen05.mma (M = 128, K = 32 bytes) cost as a function
// of its kind (bf16 / tf32), its width N and the number of TMEM accumulators the issuing thread rotates over?
// The GEMM designs of DESIGN.md 4.2 / 4.2b / 6 rest on this number (an M = 128 MMA costs about the same for
// N = 32 as for N = 256, so narrow tiles buy parallelism, not time); scripts/mma_cost.py sweeps it.
//
// One warp allocates the whole TMEM, one thread issues `iters` groups of four MMAs on zero-filled, 128-byte
// swizzled K-major operand tiles that stay in shared memory (no TMA, no epilogue: the tensor pipe alone) and
// commits them to an mbarrier; clock64 around the issue loop and around issue + completion.
#include <stdio.h>

#include "../../include/llama3_b200.h"
#include "gemm_tc_dev.cuh"

namespace {
constexpr int PROBE_A_BYTES = 128 * 128, PROBE_B_BYTES = 256 * 128;
constexpr int PROBE_SMEM = PROBE_A_BYTES + PROBE_B_BYTES + 1024 /*align*/ + 64 /*barrier, TMEM slot*/;

template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int nacc, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t a_addr = tiles, b_addr = tiles + PROBE_A_BYTES;
  const uint32_t bar = b_addr + PROBE_B_BYTES, slot = bar + 16;
  uint32_t* words = reinterpret_cast<uint32_t*>(smem_raw + (tiles - raw));
  for (int i = threadIdx.x; i < (PROBE_A_BYTES + PROBE_B_BYTES) / 4; i += blockDim.x) words[i] = 0u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the MMA reads the operands through the async proxy
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem_raw + (slot - raw));
  if (threadIdx.x == 0) {
    constexpr uint32_t FMT = KIND == TC_BF16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(b_addr);
    auto group = [&](int i) {  // four 32-byte K slices of the swizzle atom, as in the GEMM main loops
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t acc = tmem_base + (uint32_t)(((i * 4 + kk) % nacc) * n);
        tc_mma<KIND>(acc, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, 1u);
      }
    };
    for (int i = 0; i < 4; ++i) group(i);  // warm-up
    tc_commit(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) group(i);
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 1);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;      // issue loop alone (the queue may run ahead of the tensor pipe)
    out[2 * blockIdx.x + 1] = t2 - t0;  // until the last MMA has completed
  }
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}
}  // namespace

// kind 0 = bf16 (K = 16 per MMA), 1 = tf32 (K = 8); n = 16..256 (multiple of 16); nacc * n <= 512;
// ctas CTAs run the same loop concurrently (1 = an otherwise idle chip, 148 = every SM);
// out: cycles per MMA, [0] = issue loop, [1] = issue + completion, both the maximum over the CTAs.
extern "C" int l3_probe_mma(int device, int kind, int n, int nacc, int iters, int ctas, double* cycles_per_mma) {
  if ((kind != 0 && kind != 1) || n < 16 || n > 256 || n % 16 || nacc < 1 || nacc * n > 512 || iters < 1 || ctas < 1 ||
      ctas > 1024 || !cycles_per_mma)
    return L3_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return L3_ECUDA;
  long long* d = nullptr;
  if (cudaMalloc((void**)&d, (size_t)ctas * 2 * sizeof(long long)) != cudaSuccess) return L3_ENOMEM;
  cudaError_t e;
  if (kind == 0) {
    e = cudaFuncSetAttribute(mma_probe_kernel<TC_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM);
    if (e == cudaSuccess) mma_probe_kernel<TC_BF16><<<ctas, 128, PROBE_SMEM>>>(n, nacc, iters, d);
  } else {
    e = cudaFuncSetAttribute(mma_probe_kernel<TC_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PROBE_SMEM);
    if (e == cudaSuccess) mma_probe_kernel<TC_TF32X3><<<ctas, 128, PROBE_SMEM>>>(n, nacc, iters, d);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int rc = L3_OK;
  if (e != cudaSuccess) {
    rc = L3_ECUDA;
  } else {
    long long* h = (long long*)malloc((size_t)ctas * 2 * sizeof(long long));
    if (!h) {
      rc = L3_ENOMEM;
    } else if (cudaMemcpy(h, d, (size_t)ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = L3_ECUDA;
    } else {
      long long issue = 0, total = 0;
      for (int i = 0; i < ctas; ++i) {
        if (h[2 * i] > issue) issue = h[2 * i];
        if (h[2 * i + 1] > total) total = h[2 * i + 1];
      }
      cycles_per_mma[0] = (double)issue / (4.0 * iters);
      cycles_per_mma[1] = (double)total / (4.0 * iters);
    }
    free(h);
  }
  cudaFree(d);
  return rc;
}
