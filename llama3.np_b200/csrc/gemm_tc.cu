// Tensor-core projections for many activation rows: C[M, N] = A[M, K] * W[N, K]^T on the
// 5th-generation tensor cores - tcgen05.mma issued by one thread, operands staged in shared
// memory by TMA (128-byte swizzle), the fp32 accumulator tile in TMEM, read back with
// tcgen05.ld for the fused epilogue (store / residual / SwiGLU / RoPE + KV append).
//
// Two operand kinds:
//   TC_BF16   - bf16 A and W (bf16 mode): one kind::f16 MMA per 16-wide K slice.
//   TC_TF32X3 - fp32 mode.  A plain TF32 MMA keeps 10 mantissa bits and would break the 1e-4
//               logit bar (measured 2e-3), so every fp32 operand is stored as an exact pair
//               hi = x with the low 13 mantissa bits cleared, lo = x - hi, and each K slice
//               issues three kind::tf32 MMAs into the same accumulator:
//               A_lo*W_hi + A_hi*W_lo + A_hi*W_hi.  The dropped A_lo*W_lo term is ~2^-22
//               relative.  The tensor core rounds the accumulator toward zero once per MMA
//               instruction (measured: a single accumulator drifts by ~0.5 ulp per instruction,
//               7e-6 at K = 768), so the K slices are dealt round-robin to THREE main TMEM
//               accumulators and the two correction products go to a fourth; the epilogue adds
//               the four in fp32.  End-to-end logits stay within ~1e-6 of the float64 reference.
//
// Warp roles in a 192-thread CTA (one 128 x BN output tile per CTA):
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: epilogue
// synchronised by a STAGES-deep ring of full/empty mbarriers and one accumulator barrier.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "gemm_tc.h"

// Debug timeline (l3_debug_tc_timeline): CTA (0,0) stamps clock64 at pipeline milestones when enabled.
__device__ unsigned long long g_tc_dbg[64];
__device__ int g_tc_dbg_on = 0;
#define TC_STAMP(i)                                                             \
  do {                                                                          \
    if (g_tc_dbg_on && blockIdx.x == 0 && blockIdx.y == 0) g_tc_dbg[i] = clock64(); \
  } while (0)

#include "gemm_tc_dev.cuh"

// Persistent kernel: every CTA walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (row
// tile fastest, so the CTAs running at one time share a few weight tiles through L2).  The TMA
// producer and the MMA issuer run ahead across tile boundaries; with two TMEM accumulator buffers
// the epilogue warps drain tile i while the tensor core already works on tile i + 1.
template <int KIND, int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               int rows, int N, int K, int a_box_rows, int nst, int ksplit, float* part, int* tile_cnt, EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch();  // the next kernel may start its own prologue now
  if (threadIdx.x == 0) {
    TC_STAMP(0);
    // descriptor fetches overlap the barrier / TMEM set-up (the first TMA otherwise waits ~0.7 us for them)
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB0));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1));
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1));
  }
  TcPipe p;
  tc_pipe_setup<KIND, BN>(p, smem_raw, nst);
  pdl_wait();  // everything above overlapped the previous kernel; its outputs are visible from here
  if (threadIdx.x == 0) TC_STAMP(1);
  tc_gemm_run<KIND, BN, EPI>(p, &tmA0, &tmA1, &tmB0, &tmB1, rows, N, K, a_box_rows, ksplit, part, tile_cnt, e, blockIdx.x, gridDim.x);
  if (threadIdx.x == 64) TC_STAMP(34);
  tc_pipe_teardown<KIND, BN>(p);
  if (threadIdx.x == 0) TC_STAMP(35);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// [rows, K] row-major matrix, box = 128 bytes of K x box_rows rows, 128-byte swizzle, zero OOB fill
static bool make_map(CUtensorMap* tm, const void* ptr, bool is_bf16, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const size_t es = is_bf16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct MapCache {
  std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> m;
  std::mutex mu;
  const CUtensorMap* get(const void* ptr, bool is_bf16, int rows, int K, int box_rows) {
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_tuple(ptr, (int)is_bf16, rows, K, box_rows);
    auto it = m.find(key);
    if (it != m.end()) return &it->second;
    CUtensorMap tm;
    if (!make_map(&tm, ptr, is_bf16, rows, K, box_rows)) return nullptr;
    return &m.emplace(key, tm).first->second;
  }
};
static MapCache g_maps;

const CUtensorMap* tc_get_map(const void* ptr, bool is_bf16, int rows, int cols, int box_rows) {
  return g_maps.get(ptr, is_bf16, rows, cols, box_rows);
}

int tc_debug_timeline(int enable, unsigned long long* out64) {
  if (out64 && cudaMemcpyFromSymbol(out64, g_tc_dbg, sizeof(unsigned long long) * 64) != cudaSuccess) return -1;
  unsigned long long z[64] = {0};
  cudaMemcpyToSymbol(g_tc_dbg, z, sizeof z);
  return cudaMemcpyToSymbol(g_tc_dbg_on, &enable, sizeof(int)) == cudaSuccess ? 0 : -1;
}

void tc_forget_maps() {
  std::lock_guard<std::mutex> g(g_maps.mu);
  g_maps.m.clear();
}

bool tc_gemm_supported(int K) { return K % 8 == 0 && encode_fn() != nullptr; }

int tc_pick_bn(int kind, int rows, int N) {
  const int tm = (rows + 127) / 128;
  const int cand[4] = {256, 128, 64, 32};
  // 3xTF32 keeps four accumulators per tile: BN = 128 fills the TMEM with ONE buffer (no overlap of
  // epilogue and mainloop), BN = 64 leaves room for two - preferred unless L3_TF32_BN128 is set
  static const bool tf32_128 = getenv("L3_TF32_BN128") && atoi(getenv("L3_TF32_BN128")) != 0;
  if (kind == TC_TF32X3_2) {
    // An M = 128 tcgen05.mma occupies the tensor pipe for 128 cycles whatever N is (profiles/r02_mma_cost.jsonl:
    // 132.7 cycles per MMA for N = 16 .. 256, bf16 and tf32 alike), so a tile's main loop costs the same for every
    // width and only the number of tile rounds matters: always the widest tile.
    static const int lm_bn = getenv("L3_LM_BN") ? atoi(getenv("L3_LM_BN")) : 256;
    return lm_bn;
  }
  for (int i = (kind == TC_TF32X3 ? (tf32_128 ? 1 : 2) : 0); i < 4; ++i) {
    const int bn = cand[i];
    if ((long)tm * ((N + bn - 1) / bn) >= 120 || bn == 32) return bn;
  }
  return 32;
}

template <int KIND, int BN, int EPI>
static cudaError_t launch_tc_t(const TcGemmArgs& a, cudaStream_t s) {
  using Cf = TcCfg<KIND, BN>;
  auto kern = gemm_tc_kernel<KIND, BN, EPI>;
  static bool attr_done[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 15]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM);
    if (e != cudaSuccess) return e;
    attr_done[dev & 15] = true;
  }
  constexpr bool b16 = KIND == TC_BF16;
  const int a_box = a.rows >= 128 ? 128 : ((a.rows + 7) & ~7);
  const CUtensorMap* A0 = g_maps.get(a.A[0], b16, a.rows, a.K, a_box);
  const CUtensorMap* B0 = g_maps.get(a.W[0], b16, a.N, a.K, BN);
  const CUtensorMap* A1 = Cf::PARTS == 2 ? g_maps.get(a.A[1], b16, a.rows, a.K, a_box) : A0;
  const CUtensorMap* B1 = Cf::PARTS == 2 ? g_maps.get(a.W[1], b16, a.N, a.K, BN) : B0;
  if (!A0 || !B0 || !A1 || !B1) return cudaErrorInvalidValue;
  const int ntiles = ((a.N + BN - 1) / BN) * ((a.rows + 127) / 128);
  static int n_sm[16] = {0};
  if (!n_sm[dev & 15]) cudaDeviceGetAttribute(&n_sm[dev & 15], cudaDevAttrMultiProcessorCount, dev);
  const int sms = n_sm[dev & 15] > 0 ? n_sm[dev & 15] : 148;
  const int nkb_all = (a.K + Cf::BK - 1) / Cf::BK;
  // K-split (deterministic, see the epilogue) when the output tiles cover less than half the machine:
  // short-and-wide decode projections (N = D = 288: 18 tiles) otherwise leave most SMs idle while
  // each CTA walks the whole K loop.  Needs caller-owned scratch (TcGemmArgs::part / tile_cnt).
  int ksplit = 1;
  if (a.part && a.tile_cnt && ntiles * 2 <= sms && ntiles <= a.tile_cnt_len) {
    ksplit = std::min(std::min(8, sms / ntiles), nkb_all / 3);
    const size_t need = (size_t)((a.rows + 127) / 128 * 128) * ((a.N + BN - 1) / BN * BN) * sizeof(float);
    while (ksplit > 1 && need * ksplit > a.part_bytes) --ksplit;
    // every slice must own at least one k-block
    while (ksplit > 1 && (ksplit - 1) * ((nkb_all + ksplit - 1) / ksplit) >= nkb_all) --ksplit;
    if (ksplit < 1) ksplit = 1;
  }
  dim3 grid(std::min(ntiles * ksplit, sms));
  // ring depth: no deeper than the K loop; L3_GEMM_MAXSTAGES caps it further so that two kernels' CTAs
  // fit one SM (programmatic dependent launch can then overlap a kernel's prologue with its predecessor)
  static const int max_st = getenv("L3_GEMM_MAXSTAGES") ? atoi(getenv("L3_GEMM_MAXSTAGES")) : Cf::STAGES;
  const int nkb = (nkb_all + ksplit - 1) / ksplit;
  const int nst = std::max(2, std::min(std::min(Cf::STAGES, max_st), nkb));
  const size_t smem = (size_t)nst * Cf::STAGE_BYTES + Cf::EPI_BYTES + 1024 + 512;
  return launch_k(kern, grid, dim3(192), smem, s, *A0, *A1, *B0, *B1, a.rows, a.N, a.K, a_box, nst, ksplit, a.part, a.tile_cnt, a.e);
}


template <int KIND, int BN>
static cudaError_t launch_tc_e(const TcGemmArgs& a, cudaStream_t s) {
  switch (a.epi) {
    case EPI_STORE: return launch_tc_t<KIND, BN, EPI_STORE>(a, s);
    case EPI_RESID: return launch_tc_t<KIND, BN, EPI_RESID>(a, s);
    case EPI_SWIGLU: return launch_tc_t<KIND, BN, EPI_SWIGLU>(a, s);
    case EPI_ARGMAX: return launch_tc_t<KIND, BN, EPI_ARGMAX>(a, s);
    default: return launch_tc_t<KIND, BN, EPI_ROPE_KV>(a, s);
  }
}

cudaError_t launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s) {
  const int bn = a.bn > 0 ? a.bn : tc_pick_bn(a.kind, a.rows, a.N);
  if (a.kind == TC_BF16) {
    switch (bn) {
      case 256: return launch_tc_e<TC_BF16, 256>(a, s);
      case 128: return launch_tc_e<TC_BF16, 128>(a, s);
      case 64: return launch_tc_e<TC_BF16, 64>(a, s);
      default: return launch_tc_e<TC_BF16, 32>(a, s);
    }
  }
  if (a.kind == TC_TF32X3_2) {  // one main accumulator: only the LM-head epilogues are instantiated
    const bool argmax = a.epi == EPI_ARGMAX;
    switch (bn) {
      case 256: return argmax ? launch_tc_t<TC_TF32X3_2, 256, EPI_ARGMAX>(a, s) : launch_tc_t<TC_TF32X3_2, 256, EPI_STORE>(a, s);
      case 128: return argmax ? launch_tc_t<TC_TF32X3_2, 128, EPI_ARGMAX>(a, s) : launch_tc_t<TC_TF32X3_2, 128, EPI_STORE>(a, s);
      default: return argmax ? launch_tc_t<TC_TF32X3_2, 64, EPI_ARGMAX>(a, s) : launch_tc_t<TC_TF32X3_2, 64, EPI_STORE>(a, s);
    }
  }
  switch (bn) {
    case 128: return launch_tc_e<TC_TF32X3, 128>(a, s);
    case 64: return launch_tc_e<TC_TF32X3, 64>(a, s);
    default: return launch_tc_e<TC_TF32X3, 32>(a, s);
  }
}

// ------------------------------------------------------------------------------ operand preparation
// fp32 -> exact (hi, lo) pair for the 3xTF32 scheme; used for weights at load time.
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h, l;
    split_tf32(src[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
cudaError_t launch_split_tf32(const float* src, float* hi, float* lo, int64_t n, cudaStream_t s) {
  int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  if (grid < 1) grid = 1;
  split_tf32_kernel<<<grid, 256, 0, s>>>(src, hi, lo, n);
  return cudaGetLastError();
}
