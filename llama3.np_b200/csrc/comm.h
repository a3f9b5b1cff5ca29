// Tensor-parallel communication state and collectives (comm.cu).
#pragma once
#include <stdint.h>

#define L3_MAX_TP 8
#define L3_ONESHOT_MAX_FLOATS (64 * 1024)  // 256 KB per sender slot: up to 16 rows of 4096 fp32

struct L3Model;

struct L3Comm {
  int rank = 0, world = 1;
  void* nccl = nullptr;               // ncclComm_t
  bool oneshot = false;               // peer-memory one-shot all-reduce available
  int slot_floats = 0;
  void* area = nullptr;               // this rank's receive area (slots | flags | epoch)
  void* peer_base[L3_MAX_TP] = {};    // every rank's area as mapped in this process
};

void tp_destroy(L3Model* m);
int tp_barrier(L3Model* m);  // NCCL barrier on the model's stream (no-op without a communicator)
int tp_allreduce_sum(L3Model* m, const float* src, float* dst, int64_t count);
int tp_allreduce_sum_bf16(L3Model* m, void* buf, int64_t count);
int tp_allreduce_max_u64(L3Model* m, unsigned long long* keys, int count);
int tp_allgather(L3Model* m, const float* send, float* recv, int64_t count);
