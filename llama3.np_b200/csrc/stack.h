// Interface of the cluster-resident batched-decode kernel (decode_stack.cu): every layer of one decode
// step (Llama.__call__ with L == 1, llama3.py:285-304) for a block of sequences inside one thread-block cluster.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct StackLayer {
  const float* wpack;      // [C][cta_floats] per-CTA slabs of this layer's four projections (stack_pack_layer)
  const float* norm_in;    // [D]
  const float* norm_post;  // [D]
  float* ck;               // [maxB, KVHN, M, HD] K cache (post-RoPE), fp32
  float* cv;
};

struct StackArgs {
  const StackLayer* layers;  // device array [NL]
  int NL, B, M;
  const float* embed;        // [vocab, D] fp32
  const float* norm_final;   // [D]
  float eps;
  const float* cos_tab;      // [M, HD / 2]
  const float* sin_tab;
  const int* scal;           // [1] steps done so far, [2] position base: this step runs at pos = scal[2] + scal[1] + 1
  const int32_t* d_next;     // [B] input token of every sequence
  float* xlast_hi;           // [B, D] final-normed rows as exact TF32 (hi, lo) pairs: the LM head's operands
  float* xlast_lo;
  int pf_rows;               // cache rows per (sequence, head) of the NEXT layer requested into L2 during the projection phases (0 = off)
  int kv_evict_first;        // ring copies of cached K / V carry an L2 evict-first hint
  float* dbg_x;              // optional [NL][B][D]: the residual stream after every layer (null = off)
  unsigned long long* dbg;   // optional timeline [grid][64] of %globaltimer stamps (null = off)
};

// Shapes the kernel is instantiated for (fp32 weights and caches, n_heads == n_kv_heads == cluster size).
bool decode_stack_supported(int D, int HN, int KVHN, int HD, int FD, int M);
// bytes of one layer's packed copy, and the pack itself (device to device, from the [out, in] matrices of model.h)
size_t decode_stack_pack_bytes(int D, int HN, int HD, int FD);
cudaError_t decode_stack_pack_layer(const float* wqkv, const float* wo, const float* w13, const float* w2, int D, int HN,
                                    int HD, int FD, float* wpack, cudaStream_t s);
// sequences per cluster and clusters that can be co-resident on this device (0 = the kernel cannot launch here)
int decode_stack_seqs_per_cluster();
int decode_stack_max_clusters();
cudaError_t launch_decode_stack(const StackArgs& a, int D, int HN, int HD, int FD, cudaStream_t s);
// next_ids / token table / step scalars from the fused-argmax keys of the LM head that follows the stack
cudaError_t launch_stack_finalize(unsigned long long* best, int B, int32_t* next_ids, int64_t* tokens, int stride,
                                  int* scal, cudaStream_t s);
