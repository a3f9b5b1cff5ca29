// Interface of the persistent batch-1 decode kernel (decode_mega.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MG_MAX_K 14336  // longest activation vector staged in shared memory (8B-shaped FFN width)

struct MegaLayer {
  const void *wqkv, *wo, *w13, *w2;
  const float *norm_in, *norm_post;
  void *ck, *cv;
};

struct MegaArgs {
  const MegaLayer* layers;  // device array [NL]
  int NL, D, HN, KVHN, HD, FD, VS, M;
  const void* embed;
  const void* lm_head;
  const float* norm_final;
  float eps;
  const float *cos_tab, *sin_tab;
  float *x, *q, *ctx, *h;            // activations between phases (L2 resident)
  float *part_o, *part_ml;           // split-KV partials
  int* attn_cnt;
  int nsplit;
  int* scal;                         // [0] pos [1] step [2] prompt length
  int32_t* d_next;                   // [1] previous token in, next token out
  int64_t* d_tokens;                 // row 0 of the [maxB, M] token table
  unsigned long long* d_best;        // packed (value, index) argmax key, zero between steps
  unsigned *bar_cnt, *bar_gen;       // grid barrier state
  // Flag-in-data hand-offs: the residual stream after the two row-parallel projections and the attention output
  // travel as 8-byte {value, exchange number} words that the consuming phase polls - no grid barrier there.  With
  // tensor parallelism (tp_world > 1) the residual words go to EVERY rank's region through peer memory and the
  // consumer sums them in rank order; on one GPU the region is local and holds one sender.
  int tp_rank, tp_world, ll_words;
  unsigned long long* peer_ll[8];    // every rank's receive region [2][world][ll_words] of {value, epoch} words, as mapped here
  unsigned long long* ll_ctx;        // local: tagged attention output [HN * HD]
  unsigned long long* ll_xsum;       // local: [2][L3_LL_VEC] the residual stream summed over ranks (tp_world >= 4: two-level sum)
  unsigned* epoch;                   // local count of exchanges so far (carried from launch to launch)
  unsigned long long* dbg;           // optional timeline [grid][512] of %globaltimer stamps (null = off)
};

bool decode_mega_supported(int D, int HN, int KVHN, int HD, int FD, int VS);
cudaError_t launch_decode_mega(const MegaArgs& a, bool w_bf16, int grid, cudaStream_t s);
