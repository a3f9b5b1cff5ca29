// Internal model state behind the opaque L3Model handle (include/llama3_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/llama3_b200.h"
#include "comm.h"

struct L3Layer {
  void* wqkv = nullptr;  // [(HN + 2 KVHN) * HD, D]  fused q | k | v rows       (llama3.py:166-168)
  void* wo = nullptr;    // [D, HN * HD]                                          (llama3.py:211)
  void* w13 = nullptr;   // [2 FD, D]  rows interleaved gate_j, up_j              (llama3.py:99-100)
  void* w2 = nullptr;    // [D, FD]                                               (llama3.py:102)
  float* norm_in = nullptr;
  float* norm_post = nullptr;
  // fp32 mode, tensor-core path: exact TF32 (hi, lo) copies of wqkv, wo, w13, w2 (gemm_tc.cu)
  float* w_hi[4] = {nullptr, nullptr, nullptr, nullptr};
  float* w_lo[4] = {nullptr, nullptr, nullptr, nullptr};
  void* ck = nullptr;    // [maxB, KVHN, M, HD]  K cached post-RoPE               (llama3.py:138-153, 184)
  void* cv = nullptr;    // [maxB, KVHN, M, HD]
};

struct L3Graph {
  int B = 0;
  int ragged = 0;  // 0 uniform positions; 1 / 2: per-sequence positions with offset 0 / -1
  bool warmed = false;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t nodes = 0;
};

struct L3Model {
  L3Config cfg{};
  bool bf16 = false;
  // local (per tensor-parallel rank) dimensions
  int D = 0, HD = 0, G = 1, HN = 0, KVHN = 0, FD = 0, VS = 0, M = 0, maxB = 0, qkv_rows = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // weights
  void* embed = nullptr;    // [vocab, D] (replicated)
  void* lm_head = nullptr;  // [VS, D] (vocab-sharded under TP)
  float* norm_final = nullptr;
  std::vector<L3Layer> layers;
  float* cos_tab = nullptr;  // [M, HD/2] fp32 copies of the host float64 tables (llama3.py:31-38)
  float* sin_tab = nullptr;
  std::vector<char> loaded;
  bool rope_set = false, finalized = false;
  float* stage = nullptr;
  // activation workspace for one chunk of cap_tok rows
  int cap_tok = 0, max_split = 1;
  float *x = nullptr, *xn = nullptr, *q = nullptr, *ctx = nullptr, *h = nullptr, *xlast = nullptr, *logits = nullptr;
  float *part_o = nullptr, *part_ml = nullptr;
  int* attn_cnt = nullptr;  // [maxB * HN] arrival counters of the split-KV decode attention
  // tensor-core GEMM operands: fp32 mode keeps (hi, lo) pairs - the hi part lives in xn / ctx /
  // h / xlast themselves - bf16 mode keeps bf16 mirrors
  bool tc_ok = false;
  float *xn_lo = nullptr, *ctx_lo = nullptr, *h_lo = nullptr, *xlast_lo = nullptr;
  float *lm_hi = nullptr, *lm_lo = nullptr;
  void *xn16 = nullptr, *ctx16 = nullptr, *h16 = nullptr, *xlast16 = nullptr, *q16 = nullptr;
  float* gemm_part = nullptr;  // K-split scratch of the tensor-core GEMMs (32 MB) + its tile counters
  int* gemm_cnt = nullptr;
  bool attn_tc_ok = false;  // bf16 tensor-core prefill attention (attention_tc.cu)
  int32_t* d_ids = nullptr;   // [maxB, M] staged prompt
  int32_t* d_next = nullptr;  // [maxB] argmax of the last step = input of the next
  int32_t* d_fwd_ids = nullptr;   // [maxB, M] ids of l3_forward (a prompt pending in d_ids for generate is left alone)
  int32_t* d_fwd_next = nullptr;  // [maxB] int32 argmax of l3_forward / l3_forward_dev (never the generate loop's input)
  int* d_scal = nullptr;      // [0] start_pos  [1] output column  [2] prompt length  [3] zero
  int64_t* d_tokens = nullptr;  // [maxB, M] generated ids, column = step
  int64_t* d_fwd_arg = nullptr; // [maxB] argmax of l3_forward
  unsigned long long* d_best = nullptr;  // [maxB] fused-argmax keys (gemm_tc.cu EPI_ARGMAX)
  int32_t* h_next = nullptr;  // pinned
  // greedy-loop state
  int gen_B = 0, gen_L = 0, gen_step = 0, pend_B = 0, pend_L = 0;
  int gen_off = 0, pend_off = 0;
  // ragged batches (l3_generate_ragged): per-sequence prompt length, current position, last prompt row, EOS flag
  int *d_rowlen = nullptr, *d_rowpos = nullptr, *d_done = nullptr;
  int32_t* d_lastrow = nullptr;  // decode position base offset: 0 = llama3.py schedule, -1 = llama3_simple.py
  std::vector<L3Graph> graphs;
  // measurement
  int64_t launch_acc = 0;
  void* l2buf = nullptr;
  int l2_phase = 0;
  // persistent batch-1 decode kernel (decode_mega.cu)
  bool mega_ok = false;
  int n_sm = 0;
  void* d_mega_layers = nullptr;
  unsigned* d_mega_bar = nullptr;  // [0] arrival count [1] generation [2] exchange counter (single GPU)
  unsigned long long* d_mega_ll = nullptr;  // tagged words: [2][L3_LL_WORDS] residual stream (single GPU) | [HN * HD] attention output | [2][L3_LL_VEC] summed stream
  unsigned long long* d_mega_dbg = nullptr;  // L3_MEGA_DBG=1: [n_sm][512] timeline stamps of the last step
  // cluster-resident batched decode (decode_stack.cu): per-layer packed weight slabs + pointer table
  bool stack_ok = false;
  int stack_min_B = 0;
  void* d_stack_layers = nullptr;
  std::vector<float*> stack_wpack;
  float* d_stack_dbgx = nullptr;              // L3_STACK_DBG=1: [NL][maxB][D] residual stream after every layer
  unsigned long long* d_stack_dbg = nullptr;  // L3_STACK_DBG=1: [grid][64] timeline stamps of the last step
  // tensor parallel (comm.cu)
  L3Comm* comm = nullptr;
  float* logits_loc = nullptr;   // [maxB, VS] local vocabulary slice (G > 1)
  float* logits_all = nullptr;   // [G, maxB, VS] all-gathered slices (G > 1)
  char err[512] = "";
};
