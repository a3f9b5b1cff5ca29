// Tensor-parallel communication state and collectives (comm.cu).
#pragma once
#include <stdint.h>

#define L3_MAX_TP 8

// Flag-in-data ("LL") receive region used INSIDE the persistent decode kernel: 8-byte words {fp32 value, epoch}
// written with single 8-byte stores over NVLink, so the receiver polls the data itself - no system fence, no flag
// store, no barrier between the producing epilogue and the consuming phase.  Per (buffer, sender): a vector of up
// to L3_LL_VEC values, then two words for the packed argmax key.
#define L3_LL_VEC 16384
#define L3_LL_WORDS (L3_LL_VEC + 8)
// The same protocol as a stand-alone multi-CTA all-reduce kernel (allreduce_ll_kernel) for messages between kernels:
// up to L3_LL2_WORDS fp32 values (2 MB of payload, e.g. 128 rows of 4096) per call - every decode-sized sum over ranks,
// so that no NCCL call remains in a decode step.  Area: 2 x world x 4 MB.
#define L3_LL2_WORDS (512 * 1024)

struct L3Model;

struct L3Comm {
  int rank = 0, world = 1;
  void* nccl = nullptr;               // ncclComm_t
  bool oneshot = false;               // every rank mapped every rank's receive area: sums over ranks run through peer memory
  void* area = nullptr;               // this rank's receive area (counters | decode_mega region | all-reduce region)
  void* peer_base[L3_MAX_TP] = {};    // every rank's area as mapped in this process
};

void tp_destroy(L3Model* m);
// byte offsets inside a rank's receive area (same on every rank of a communicator): 64 bytes of counters ([0] unused,
// [1] exchanges of decode_mega_kernel so far, [2] all-reduce calls so far, [3] CTAs of the running all-reduce that have
// finished), then the two flag-in-data regions
static inline size_t tp_epoch_off() { return 0; }
static inline size_t tp_ll_off() { return 64; }                                                       // [2][world][L3_LL_WORDS] x 8 bytes
static inline size_t tp_ll2_off(int world) { return tp_ll_off() + (size_t)2 * world * L3_LL_WORDS * 8; }   // [2][world][L3_LL2_WORDS] x 8 bytes
static inline size_t tp_area_bytes(int world) { return tp_ll2_off(world) + (size_t)2 * world * L3_LL2_WORDS * 8; }
int tp_barrier(L3Model* m);  // NCCL barrier on the model's stream (no-op without a communicator)
int tp_allreduce_sum(L3Model* m, const float* src, float* dst, int64_t count);
int tp_allreduce_sum_bf16(L3Model* m, void* buf, int64_t count);
int tp_allreduce_max_u64(L3Model* m, unsigned long long* keys, int count);
int tp_allgather(L3Model* m, const float* send, float* recv, int64_t count);
