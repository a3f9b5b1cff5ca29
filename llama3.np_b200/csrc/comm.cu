// Tensor-parallel communication for the 8B-shaped configs: one process per GPU,
// heads / FFN columns / vocabulary rows sharded at load time (place_of, l3_api.cu), and one
// sum over ranks after the attention output projection and after the FFN down projection
// (the two row-parallel matrices).  The reference has no counterpart (single process, NumPy).
//
// Two transports:
//   * flag-in-data peer-memory all-reduce (decode-sized messages, up to 2 MB): every rank writes its partial
//     vector straight into a region of every peer's receive area with NVLink P2P stores as 8-byte
//     {value, call number} words; the receiver polls the data itself and sums the `world` regions in rank
//     order - identical arithmetic on every rank, so the replicated residual stream never diverges; no
//     system fence, no flag, any number of CTAs (allreduce_ll_kernel; decode_mega_kernel runs the same
//     protocol inside the kernel).  The areas come from cudaMalloc and travel between the processes as
//     CUDA IPC handles.  NVSwitch gives every pair full bandwidth, so a flat exchange is the
//     latency-optimal schedule for these messages.  Measured at TP 2, 8B batch 32: 4.62 ms per decode step
//     against 5.14 ms with 64 ncclAllReduce calls per token.
//   * NCCL (prefill-sized messages, bootstrap, u64-max for the vocabulary-sharded argmax).
//     libnccl.so.2 is resolved with dlopen at l3_tp_init time - the library the process already
//     holds (torch's) is reused, single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "comm.h"
#include "model.h"

namespace {
struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  char why[256] = "";
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.so ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("L3_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.so) break;
    snprintf(api.why, sizeof api.why, "%s", dlerror());
  }
  if (!api.so) return nullptr;
#define SYM(field, name)                                            \
  *(void**)(&api.field) = dlsym(api.so, name);                      \
  if (!api.field) {                                                 \
    snprintf(api.why, sizeof api.why, "missing symbol %s", name);   \
    api.so = nullptr;                                               \
    return nullptr;                                                 \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return &api;
}

void comm_err(L3Model* m, const char* what, const char* detail) {
  if (m) snprintf(m->err, sizeof m->err, "%s: %s", what, detail ? detail : "");
}
}  // namespace

// ------------------------------------------------------------------------------ flag-in-data all-reduce
// dst = sum over ranks of src for up to L3_LL2_WORDS values, any number of CTAs, no fence and no flag: every value
// travels as an 8-byte {fp32 bits, call number} word (single-copy atomic, also over NVLink) into region
// [call & 1][sender] of every rank; each thread then polls the words of ITS elements from every sender until they
// carry this call's number and adds them in rank order (identical bits on every rank).  The call number lives on the
// device (the kernel is captured in CUDA graphs); the last CTA to finish advances it.  Two buffers suffice: a rank
// starts call e + 2 only after it has received every peer's part of e + 1, and a peer sends e + 1 after reading e.
struct LLArgs {
  unsigned long long* peer[L3_MAX_TP];  // every rank's region [2][world][L3_LL2_WORDS], as mapped here
  uint32_t* epoch;                      // local: [0] calls so far, [1] CTAs of the running call that have finished
  const float* src;
  float* dst;
  int count, rank, world;
  int two_phase;                        // reduce-scatter + all-gather instead of all-to-all (see the kernel)
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void ll_store4(unsigned long long* d, float4 v, uint32_t epoch) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(__float_as_uint(v.x)), "r"(epoch),
               "r"(__float_as_uint(v.y)), "r"(epoch) : "memory");
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(d + 2), "r"(__float_as_uint(v.z)), "r"(epoch),
               "r"(__float_as_uint(v.w)), "r"(epoch) : "memory");
}
// four tagged values; a lost peer must fail the launch, not hang the GPU - but ranks may legitimately arrive seconds
// apart, so the limit is wall time
__device__ __forceinline__ float4 ll_wait4(const unsigned long long* w, uint32_t epoch, unsigned long long timeout_ns, unsigned long long& t0) {
  uint4 lo, hi;
  uint32_t spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(w) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(w + 2) : "memory");
    if (lo.y == epoch && lo.w == epoch && hi.y == epoch && hi.w == epoch) break;
    if ((++spins & 0x3ff) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (!t0) t0 = now;
      if (now - t0 > timeout_ns) __trap();
    }
  }
  return make_float4(__uint_as_float(lo.x), __uint_as_float(lo.z), __uint_as_float(hi.x), __uint_as_float(hi.z));
}

// all-to-all (small messages: one NVLink hop; every rank receives world x count words and sums them itself), or, when
// world x count is large, reduce-scatter + all-gather in ONE kernel without any grid synchronisation: rank p owns the
// p-th slice of the vector; everybody sends its slice-p partials to p, p sums them in rank order and sends the sums to
// everybody; each step is "poll the words this thread needs".  Two hops, but 2 x count words per rank on the wire and in
// the polls instead of world x count.  Either way every rank ends with bit-identical sums.
__global__ void __launch_bounds__(256) allreduce_ll_kernel(LLArgs a) {
  pdl_launch();
  pdl_wait();
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(a.epoch) + 1;  // first call writes 1: the region starts zeroed
  const size_t buf = (size_t)(epoch & 1) * a.world * L3_LL2_WORDS;
  const int n4 = a.count >> 2, stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long t0 = 0;
  if (!a.two_phase) {
    const size_t mine = buf + (size_t)a.rank * L3_LL2_WORDS;
    for (int g = t; g < n4; g += stride) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(a.src) + g);
      for (int p = 0; p < a.world; ++p) ll_store4(a.peer[p] + mine + (size_t)g * 4, v, epoch);
    }
    const unsigned long long* region = a.peer[a.rank] + buf;
    for (int g = t; g < n4; g += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < a.world; ++r) {
        const float4 v = ll_wait4(region + (size_t)r * L3_LL2_WORDS + (size_t)g * 4, epoch, a.timeout_ns, t0);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      reinterpret_cast<float4*>(a.dst)[g] = s;
    }
  } else {
    // words of a buffer: [world senders][slice] partials for my slice, then [count] the summed vector
    const int s4 = n4 / a.world;                       // float4 groups per slice (count % (4 world) == 0)
    const size_t slice_w = (size_t)s4 * 4, sums = buf + (size_t)a.world * slice_w;
    for (int g = t; g < n4; g += stride) {
      const int p = g / s4, j = g - p * s4;
      ll_store4(a.peer[p] + buf + (size_t)a.rank * slice_w + (size_t)j * 4, __ldcg(reinterpret_cast<const float4*>(a.src) + g), epoch);
    }
    const unsigned long long* mine = a.peer[a.rank] + buf;
    for (int j = t; j < s4; j += stride) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < a.world; ++r) {
        const float4 v = ll_wait4(mine + (size_t)r * slice_w + (size_t)j * 4, epoch, a.timeout_ns, t0);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      const size_t at = sums + ((size_t)a.rank * s4 + j) * 4;
      for (int p = 0; p < a.world; ++p) ll_store4(a.peer[p] + at, s, epoch);
    }
    for (int g = t; g < n4; g += stride)
      reinterpret_cast<float4*>(a.dst)[g] = ll_wait4(a.peer[a.rank] + sums + (size_t)g * 4, epoch, a.timeout_ns, t0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // every CTA has read the call number by the time the last one is done
    if (atomicAdd(a.epoch + 1, 1u) == gridDim.x - 1) {
      a.epoch[1] = 0;
      __threadfence();
      a.epoch[0] = epoch;
    }
  }
}

unsigned long long tp_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* v = getenv("L3_TP_TIMEOUT_MS");
    const long ms = v ? atol(v) : 60000;
    return (unsigned long long)(ms > 0 ? ms : 60000) * 1000000ull;
  }();
  return ns;
}

// ------------------------------------------------------------------------------ setup / teardown
extern "C" int l3_nccl_unique_id(void* out_128) {
  NcclApi* n = nccl_api();
  if (!n) return L3_ENCCL;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (n->GetUniqueId(&id) != ncclSuccess) return L3_ENCCL;
  memcpy(out_128, &id, 128);
  return L3_OK;
}

extern "C" int l3_tp_init(L3Model* m, const void* nccl_unique_id_128) {
  if (!m) return L3_EINVAL;
  if (m->G == 1) return L3_OK;
  if (m->comm) { comm_err(m, "l3_tp_init", "already initialised"); return L3_ESTATE; }
  NcclApi* n = nccl_api();
  if (!n) { comm_err(m, "libnccl.so.2 could not be loaded", "set L3_NCCL_LIB"); return L3_ENCCL; }
  if (cudaSetDevice(m->cfg.device) != cudaSuccess) return L3_ECUDA;
  L3Comm* c = new L3Comm();
  c->rank = m->cfg.tp_rank;
  c->world = m->G;
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id_128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = n->CommInitRank(&comm, c->world, id, c->rank);
  if (r != ncclSuccess) { comm_err(m, "ncclCommInitRank", n->GetErrorString(r)); delete c; return L3_ENCCL; }
  c->nccl = comm;
  m->comm = c;

  // ---- peer-memory receive area of the flag-in-data exchanges, exchanged as CUDA IPC handles.
  // Every rank walks the SAME sequence of collectives below whatever fails locally (a rank that returned early
  // would leave the others blocked in the next one), and the decision to use peer memory is the minimum
  // over ranks of "I mapped every peer": either all ranks use peer stores or all use NCCL.
  const char* off = getenv("L3_TP_ONESHOT");
  if (off && atoi(off) == 0) return L3_OK;  // an environment switch: the same on every rank of a job
  const size_t area = tp_area_bytes(c->world);
  int ok = 1;
  const char* what = "";
  cudaError_t e = cudaMalloc(&c->area, area);
  if (e != cudaSuccess) { ok = 0; what = "cudaMalloc(receive area)"; c->area = nullptr; cudaGetLastError(); }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof mine);
  if (ok) {
    cudaMemset(c->area, 0, area);
    e = cudaIpcGetMemHandle(&mine, c->area);
    if (e != cudaSuccess) { ok = 0; what = "cudaIpcGetMemHandle"; cudaGetLastError(); }
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  char *d_send = nullptr, *d_recv = nullptr;
  int* d_ok = nullptr;
  if (cudaMalloc((void**)&d_send, 64) != cudaSuccess || cudaMalloc((void**)&d_recv, 64 * c->world) != cudaSuccess ||
      cudaMalloc((void**)&d_ok, 4) != cudaSuccess) {
    comm_err(m, "cudaMalloc(setup scratch)", "out of memory");
    tp_destroy(m);
    return L3_ENOMEM;
  }
  cudaMemcpy(d_send, &mine, 64, cudaMemcpyHostToDevice);
  r = n->AllGather(d_send, d_recv, 64, ncclChar, comm, m->stream);
  cudaStreamSynchronize(m->stream);
  std::vector<cudaIpcMemHandle_t> all(c->world);
  cudaMemcpy(all.data(), d_recv, 64 * c->world, cudaMemcpyDeviceToHost);
  if (r != ncclSuccess) {
    comm_err(m, "ncclAllGather(ipc handles)", n->GetErrorString(r));
    cudaFree(d_send); cudaFree(d_recv); cudaFree(d_ok);
    tp_destroy(m);
    return L3_ENCCL;
  }
  // first agreement: did everybody export a handle?  (opening a zeroed handle would only produce noise)
  auto all_ok = [&](int mine_ok) -> int {
    cudaMemcpy(d_ok, &mine_ok, 4, cudaMemcpyHostToDevice);
    ncclResult_t rr = n->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, m->stream);
    cudaStreamSynchronize(m->stream);
    int v = 0;
    cudaMemcpy(&v, d_ok, 4, cudaMemcpyDeviceToHost);
    return rr == ncclSuccess ? v : -1;
  };
  int agreed = all_ok(ok);
  if (agreed == 1) {
    for (int p = 0; p < c->world && ok; ++p) {
      void* base = c->area;
      if (p != c->rank) {
        e = cudaIpcOpenMemHandle(&base, all[p], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {  // no peer access between these two processes
          cudaGetLastError();
          ok = 0;
          what = "cudaIpcOpenMemHandle";
          break;
        }
      }
      c->peer_base[p] = base;
    }
    // second agreement doubles as the barrier "every rank has mapped every area before anyone pushes into it"
    agreed = all_ok(ok);
  }
  cudaFree(d_send); cudaFree(d_recv); cudaFree(d_ok);
  if (agreed < 0) {
    comm_err(m, "ncclAllReduce(peer-memory agreement)", "failed");
    tp_destroy(m);
    return L3_ENCCL;
  }
  if (agreed == 1) {
    c->oneshot = true;
  } else {  // somebody could not: NCCL carries everything, on every rank
    if (!ok) fprintf(stderr, "llama3_b200: rank %d: %s failed (%s); peer-memory exchanges disabled on all ranks (NCCL carries everything)\n", c->rank, what,
                     cudaGetErrorString(e));
    for (int p = 0; p < c->world; ++p)
      if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    memset(c->peer_base, 0, sizeof c->peer_base);
    if (c->area) { cudaFree(c->area); c->area = nullptr; }
    c->oneshot = false;
  }
  return L3_OK;
}

// Barrier over the tensor-parallel group on the model's stream (end of l3_finalize: no rank starts pushing into
// peer memory, or waiting for a peer's flag, before every rank has finished loading and packing its weights).
int tp_barrier(L3Model* m) {
  L3Comm* c = m->comm;
  if (!c) return L3_OK;
  NcclApi* n = nccl_api();
  int* d = nullptr;
  if (cudaMalloc((void**)&d, 4) != cudaSuccess) return L3_ENOMEM;
  cudaMemsetAsync(d, 0, 4, m->stream);
  ncclResult_t r = n->AllReduce(d, d, 1, ncclInt, ncclSum, (ncclComm_t)c->nccl, m->stream);
  cudaStreamSynchronize(m->stream);
  cudaFree(d);
  if (r != ncclSuccess) { comm_err(m, "ncclAllReduce(barrier)", n->GetErrorString(r)); return L3_ENCCL; }
  return L3_OK;
}

void tp_destroy(L3Model* m) {
  L3Comm* c = m->comm;
  if (!c) return;
  for (int p = 0; p < c->world; ++p)
    if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->area) cudaFree(c->area);
  NcclApi* n = nccl_api();
  if (n && c->nccl) n->CommDestroy((ncclComm_t)c->nccl);
  delete c;
  m->comm = nullptr;
}

// ------------------------------------------------------------------------------ collectives
// dst = sum over ranks of src (fp32, count elements); src and dst are distinct local buffers.
int tp_allreduce_sum(L3Model* m, const float* src, float* dst, int64_t count) {
  L3Comm* c = m->comm;
  if (!c) { comm_err(m, "tensor parallel", "l3_tp_init was not called"); return L3_ESTATE; }
  if (c->oneshot && count <= L3_LL2_WORDS && (count & 3) == 0) {
    LLArgs a{};
    for (int p = 0; p < c->world; ++p)
      a.peer[p] = (unsigned long long*)((char*)c->peer_base[p] + tp_ll2_off(c->world));
    a.epoch = (uint32_t*)((char*)c->area + tp_epoch_off()) + 2;  // words 2, 3 of the counter block
    a.src = src; a.dst = dst; a.count = (int)count; a.rank = c->rank; a.world = c->world;
    a.timeout_ns = tp_timeout_ns();
    // two hops pay only when they save wire and poll volume: world x count words against 2 x count (measured at TP 2,
    // 8B batch 32: 4.62 ms per step all-to-all, 4.71 two-phase)
    static const int force2 = [] { const char* v = getenv("L3_TP_TWO_PHASE"); return v ? atoi(v) : -1; }();  // tests: 1 / 0 force the mode
    a.two_phase = c->world >= 4 && count * c->world >= 32768 && count % (4 * c->world) == 0;
    if (force2 >= 0) a.two_phase = force2 != 0 && count % (4 * c->world) == 0;
    const int n4 = (int)(count >> 2);
    const int ctas = std::max(1, std::min(64, (n4 + 255) / 256));
    cudaError_t e = launch_k(allreduce_ll_kernel, dim3(ctas), dim3(256), 0, m->stream, a);
    if (e != cudaSuccess) { comm_err(m, "allreduce_ll_kernel", cudaGetErrorString(e)); return L3_ECUDA; }
    m->launch_acc += 1;
    return L3_OK;
  }
  NcclApi* n = nccl_api();
  ncclResult_t r = n->AllReduce(src, dst, (size_t)count, ncclFloat, ncclSum, (ncclComm_t)c->nccl, m->stream);
  if (r != ncclSuccess) { comm_err(m, "ncclAllReduce", n->GetErrorString(r)); return L3_ENCCL; }
  return L3_OK;
}

// in-place sum over ranks of a bf16 buffer (prefill-sized partial projections in bf16 mode: half
// the NVLink bytes of the fp32 exchange; the residual stream itself stays fp32)
int tp_allreduce_sum_bf16(L3Model* m, void* buf, int64_t count) {
  L3Comm* c = m->comm;
  if (!c) { comm_err(m, "tensor parallel", "l3_tp_init was not called"); return L3_ESTATE; }
  NcclApi* n = nccl_api();
  ncclResult_t r = n->AllReduce(buf, buf, (size_t)count, ncclBfloat16, ncclSum, (ncclComm_t)c->nccl, m->stream);
  if (r != ncclSuccess) { comm_err(m, "ncclAllReduce(bf16)", n->GetErrorString(r)); return L3_ENCCL; }
  return L3_OK;
}

// in-place max of packed (value, index) argmax keys over ranks
int tp_allreduce_max_u64(L3Model* m, unsigned long long* keys, int count) {
  L3Comm* c = m->comm;
  if (!c) { comm_err(m, "tensor parallel", "l3_tp_init was not called"); return L3_ESTATE; }
  NcclApi* n = nccl_api();
  ncclResult_t r = n->AllReduce(keys, keys, (size_t)count, ncclUint64, ncclMax, (ncclComm_t)c->nccl, m->stream);
  if (r != ncclSuccess) { comm_err(m, "ncclAllReduce(max)", n->GetErrorString(r)); return L3_ENCCL; }
  return L3_OK;
}

// recv [world][count] <- every rank's send [count] (fp32)
int tp_allgather(L3Model* m, const float* send, float* recv, int64_t count) {
  L3Comm* c = m->comm;
  if (!c) { comm_err(m, "tensor parallel", "l3_tp_init was not called"); return L3_ESTATE; }
  NcclApi* n = nccl_api();
  ncclResult_t r = n->AllGather(send, recv, (size_t)count, ncclFloat, (ncclComm_t)c->nccl, m->stream);
  if (r != ncclSuccess) { comm_err(m, "ncclAllGather", n->GetErrorString(r)); return L3_ENCCL; }
  return L3_OK;
}
