// Interface of the persistent batched-decode kernel (decode_batch.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// One GEMM phase: tensor maps of the (hi, lo) operand pairs, shape, fused epilogue.
struct BatchGemm {
  CUtensorMap maps[4];  // A hi, A lo (activations [rows, K]), W hi, W lo (weights [N, K])
  int rows, N, K, a_box;
  EpiArgs e;
};

struct BatchLayer {
  BatchGemm qkv, wo, w13, w2;
  const float *norm_in, *norm_post;
  void *ck, *cv;
};

struct BatchArgs {
  const BatchLayer* layers;  // device array [NL]
  const BatchGemm* lm;       // LM head with the fused argmax epilogue (device)
  int NL, B, D, HN, KVHN, HD, M, nst;
  float eps;
  const void* embed;         // fp32 [vocab, D]
  const float* norm_final;
  float *x, *xn, *xn_lo, *xlast, *xlast_lo;
  AttnArgs attn;             // q, outputs (ctx hi / lo), shapes; the per-layer caches are filled in by the kernel
  int* scal;                 // [0] pos [1] step [2] position base
  int32_t* d_next;           // [B] previous tokens in, next tokens out
  int64_t* d_tokens;         // [maxB, M] token table
  unsigned long long* d_best;
  unsigned *bar_cnt, *bar_gen;
};

bool decode_batch_supported(int HD, int nrep);
int decode_batch_max_stages();
int decode_batch_bn();
cudaError_t launch_decode_batch(const BatchArgs& a, int grid, cudaStream_t s);
