// Interface of the tcgen05 tensor-core GEMM (gemm_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

// TC_TF32X3_2: the 3xTF32 scheme with ONE main accumulator instead of three (short K loops: the per-MMA
// truncation drift stays ~2e-6 up to K = 512), half the TMEM columns -> wider / double-buffered tiles
enum TcKind : int { TC_BF16 = 0, TC_TF32X3 = 1, TC_TF32X3_2 = 2 };

struct TcGemmArgs {
  int kind;             // TcKind
  const void* A[2];     // activations [rows, K]: bf16, or fp32 (hi, lo) pair
  const void* W[2];     // weights [N, K], same element kind
  int rows, N, K;
  int bn;               // 0 = choose
  int epi;
  EpiArgs e;
  // optional caller-owned scratch for the deterministic K-split of gemm_tc.cu (null: never split)
  float* part;       // [ksplit][rows padded to 128][N padded to the tile] fp32 partial tiles
  size_t part_bytes;
  int* tile_cnt;     // zero-initialised arrival counters, one per output tile
  int tile_cnt_len;
};

cudaError_t launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s);
// swapped operand roles for 9..128 activation rows (gemm_swap.cu): same arguments
bool gemm_swap_supported(int rows, int N);
cudaError_t launch_gemm_swap(const TcGemmArgs& a, cudaStream_t s);
cudaError_t launch_split_tf32(const float* src, float* hi, float* lo, int64_t n, cudaStream_t s);
bool tc_gemm_supported(int K);
int tc_pick_bn(int kind, int rows, int N);
void tc_forget_maps();
int tc_debug_timeline(int enable, unsigned long long* out64);
// cached 2-D tensor map of a row-major [rows, cols] matrix: box = 128 bytes of columns x box_rows rows,
// 128-byte swizzle (null if the driver entry point is unavailable)
const CUtensorMap* tc_get_map(const void* ptr, bool is_bf16, int rows, int cols, int box_rows);
