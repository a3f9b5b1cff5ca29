// Packed device-layout weight cache: the weight matrices exactly as l3_load_weight leaves them
// on the device - fused q|k|v rows, interleaved gate/up rows, this rank's heads / FFN columns / vocabulary rows,
// already in the model's dtype - written to one file, so that the next start streams the file straight into the
// device buffers instead of parsing the reference's .npz (llama3.py:269, utils.py:4-5), slicing, transposing and
// converting every tensor again.  Derived copies (TF32 hi/lo pairs, per-CTA slabs of the cluster-resident decode)
// are rebuilt on the device by l3_finalize: they cost milliseconds and would double the file.
//
// File: [L3PackHeader, 4096 bytes incl. the tensor table][tensor 0][tensor 1] ...   every tensor 4096-aligned.
// A file is accepted only if shape, dtype, tensor-parallel placement and the caller's digest of the SOURCE
// checkpoint all match, and every tensor's checksum verifies.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/llama3_b200.h"
#include "model.h"

void set_err(L3Model* m, const char* fmt, ...);  // l3_api.cu

namespace {

constexpr char kMagic[8] = {'L', '3', 'P', 'A', 'C', 'K', '0', '1'};
constexpr int kMaxTensors = 3 + 6 * 160;  // header table capacity is checked at save time
constexpr size_t kAlign = 4096;
constexpr size_t kChunk = (size_t)64 << 20;

struct PackTensor { uint64_t offset, bytes, checksum; };
struct L3PackHeader {
  char magic[8];
  int32_t dim, n_layers, n_heads, n_kv_heads, vocab_size, hidden_dim, dtype, tp_rank, tp_world, n_tensors;
  uint64_t header_bytes, file_bytes;
  char digest[96];  // caller's digest of the source checkpoint (hex), NUL-terminated
};

struct Item { void* ptr; size_t bytes; };

std::vector<Item> items_of(L3Model* m) {
  const size_t wb = m->bf16 ? 2 : 4;
  std::vector<Item> v;
  v.push_back({m->embed, (size_t)m->cfg.vocab_size * m->D * wb});
  v.push_back({m->lm_head, (size_t)m->VS * m->D * wb});
  v.push_back({m->norm_final, (size_t)m->D * 4});
  for (auto& L : m->layers) {
    v.push_back({L.wqkv, (size_t)m->qkv_rows * m->D * wb});
    v.push_back({L.wo, (size_t)m->D * m->HN * m->HD * wb});
    v.push_back({L.w13, (size_t)2 * m->FD * m->D * wb});
    v.push_back({L.w2, (size_t)m->D * m->FD * wb});
    v.push_back({L.norm_in, (size_t)m->D * 4});
    v.push_back({L.norm_post, (size_t)m->D * 4});
  }
  return v;
}

// order-sensitive 64-bit checksum over 8-byte words (tensor sizes are multiples of 8: dim % 8 == 0)
uint64_t checksum_update(uint64_t h, const void* p, size_t n) {
  const uint64_t* w = (const uint64_t*)p;
  for (size_t i = 0; i < n / 8; ++i) h = (h ^ w[i]) * 0x9E3779B97F4A7C15ull + (h >> 29);
  const uint8_t* t = (const uint8_t*)p + (n & ~(size_t)7);
  for (size_t i = 0; i < (n & 7); ++i) h = (h ^ t[i]) * 0x100000001B3ull;
  return h;
}

size_t header_bytes(int n_tensors) {
  const size_t raw = sizeof(L3PackHeader) + (size_t)n_tensors * sizeof(PackTensor);
  return (raw + kAlign - 1) / kAlign * kAlign;
}

void fill_header(const L3Model* m, const char* digest, int n, L3PackHeader* h) {
  memset(h, 0, sizeof *h);
  memcpy(h->magic, kMagic, 8);
  h->dim = m->cfg.dim; h->n_layers = m->cfg.n_layers; h->n_heads = m->cfg.n_heads; h->n_kv_heads = m->cfg.n_kv_heads;
  h->vocab_size = m->cfg.vocab_size; h->hidden_dim = m->cfg.hidden_dim; h->dtype = m->cfg.dtype;
  h->tp_rank = m->cfg.tp_rank; h->tp_world = m->cfg.tp_world; h->n_tensors = n;
  h->header_bytes = header_bytes(n);
  snprintf(h->digest, sizeof h->digest, "%s", digest ? digest : "");
}

struct File {
  FILE* f = nullptr;
  ~File() { if (f) fclose(f); }
};
struct Pinned {
  void* p = nullptr;
  ~Pinned() { if (p) cudaFreeHost(p); }
};

}  // namespace

#define PK_FAIL(m, code, ...) do { set_err(m, __VA_ARGS__); return code; } while (0)
#define PK_CUDA(m, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) PK_FAIL(m, L3_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); } while (0)

extern "C" int l3_save_packed(L3Model* m, const char* path, const char* source_digest) {
  if (!m || !path) return L3_EINVAL;
  for (size_t i = 0; i < m->loaded.size(); ++i)
    if (!m->loaded[i]) PK_FAIL(m, L3_ESTATE, "l3_save_packed: weight slot %zu was never loaded", i);
  PK_CUDA(m, cudaSetDevice(m->cfg.device));
  PK_CUDA(m, cudaStreamSynchronize(m->stream));
  const std::vector<Item> items = items_of(m);
  if ((int)items.size() > kMaxTensors) PK_FAIL(m, L3_EINVAL, "l3_save_packed: %zu tensors exceed the table", items.size());
  L3PackHeader h;
  fill_header(m, source_digest, (int)items.size(), &h);
  std::vector<PackTensor> tab(items.size());
  uint64_t off = h.header_bytes;
  for (size_t i = 0; i < items.size(); ++i) {
    tab[i].offset = off; tab[i].bytes = items[i].bytes; tab[i].checksum = 0;
    off += (items[i].bytes + kAlign - 1) / kAlign * kAlign;
  }
  h.file_bytes = off;
  // write beside the target and rename: a reader never sees a half-written cache
  const std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
  File f;
  f.f = fopen(tmp.c_str(), "wb");
  if (!f.f) PK_FAIL(m, L3_EINVAL, "l3_save_packed: cannot create '%s': %s", tmp.c_str(), strerror(errno));
  Pinned pin;
  PK_CUDA(m, cudaMallocHost(&pin.p, kChunk));
  std::vector<char> zeros(kAlign, 0);
  bool ok = fseek(f.f, (long)h.header_bytes, SEEK_SET) == 0;
  for (size_t i = 0; i < items.size() && ok; ++i) {
    uint64_t cs = 0x243F6A8885A308D3ull;
    for (size_t o = 0; o < items[i].bytes && ok; o += kChunk) {
      const size_t n = std::min(kChunk, items[i].bytes - o);
      PK_CUDA(m, cudaMemcpy(pin.p, (const char*)items[i].ptr + o, n, cudaMemcpyDeviceToHost));
      cs = checksum_update(cs, pin.p, n);
      ok = fwrite(pin.p, 1, n, f.f) == n;
    }
    tab[i].checksum = cs;
    const size_t pad = (kAlign - items[i].bytes % kAlign) % kAlign;
    if (ok && pad) ok = fwrite(zeros.data(), 1, pad, f.f) == pad;
  }
  if (ok) ok = fseek(f.f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof h, 1, f.f) == 1 &&
               fwrite(tab.data(), sizeof(PackTensor), tab.size(), f.f) == tab.size();
  if (ok) ok = fflush(f.f) == 0;
  fclose(f.f);
  f.f = nullptr;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    const int en = errno;
    remove(tmp.c_str());
    PK_FAIL(m, L3_EINVAL, "l3_save_packed: writing '%s' failed: %s", path, strerror(en));
  }
  return L3_OK;
}

static int read_header(L3Model* m, FILE* f, const char* path, L3PackHeader* h, std::vector<PackTensor>* tab) {
  if (fread(h, sizeof *h, 1, f) != 1 || memcmp(h->magic, kMagic, 8) != 0)
    PK_FAIL(m, L3_EINVAL, "'%s' is not a packed weight cache", path);
  if (h->n_tensors <= 0 || h->n_tensors > kMaxTensors || h->header_bytes != header_bytes(h->n_tensors))
    PK_FAIL(m, L3_EINVAL, "'%s': corrupt header", path);
  h->digest[sizeof h->digest - 1] = 0;
  tab->resize(h->n_tensors);
  if (fread(tab->data(), sizeof(PackTensor), tab->size(), f) != tab->size()) PK_FAIL(m, L3_EINVAL, "'%s': truncated table", path);
  return L3_OK;
}

extern "C" int l3_packed_info(const char* path, L3Config* cfg_out, char* digest_out, int digest_cap) {
  if (!path) return L3_EINVAL;
  File f;
  f.f = fopen(path, "rb");
  if (!f.f) PK_FAIL(nullptr, L3_EINVAL, "cannot open '%s': %s", path, strerror(errno));
  L3PackHeader h;
  std::vector<PackTensor> tab;
  const int rc = read_header(nullptr, f.f, path, &h, &tab);
  if (rc != L3_OK) return rc;
  if (cfg_out) {
    memset(cfg_out, 0, sizeof *cfg_out);
    cfg_out->dim = h.dim; cfg_out->n_layers = h.n_layers; cfg_out->n_heads = h.n_heads; cfg_out->n_kv_heads = h.n_kv_heads;
    cfg_out->vocab_size = h.vocab_size; cfg_out->hidden_dim = h.hidden_dim; cfg_out->dtype = h.dtype;
    cfg_out->tp_rank = h.tp_rank; cfg_out->tp_world = h.tp_world;
  }
  if (digest_out && digest_cap > 0) snprintf(digest_out, (size_t)digest_cap, "%s", h.digest);
  return L3_OK;
}

extern "C" int l3_load_packed(L3Model* m, const char* path, const char* source_digest) {
  if (!m || !path) return L3_EINVAL;
  if (m->finalized) PK_FAIL(m, L3_ESTATE, "l3_load_packed after l3_finalize");
  PK_CUDA(m, cudaSetDevice(m->cfg.device));
  File f;
  f.f = fopen(path, "rb");
  if (!f.f) PK_FAIL(m, L3_EINVAL, "cannot open '%s': %s", path, strerror(errno));
  L3PackHeader h, want;
  std::vector<PackTensor> tab;
  int rc = read_header(m, f.f, path, &h, &tab);
  if (rc != L3_OK) return rc;
  const std::vector<Item> items = items_of(m);
  fill_header(m, source_digest, (int)items.size(), &want);
  if (h.dim != want.dim || h.n_layers != want.n_layers || h.n_heads != want.n_heads || h.n_kv_heads != want.n_kv_heads ||
      h.vocab_size != want.vocab_size || h.hidden_dim != want.hidden_dim || h.dtype != want.dtype ||
      h.tp_rank != want.tp_rank || h.tp_world != want.tp_world || h.n_tensors != want.n_tensors)
    PK_FAIL(m, L3_EINVAL, "'%s' was packed for another shape / dtype / tensor-parallel placement", path);
  if (source_digest && strcmp(h.digest, want.digest) != 0)
    PK_FAIL(m, L3_EINVAL, "'%s' was packed from another checkpoint (digest %s, want %s)", path, h.digest, want.digest);
  for (size_t i = 0; i < items.size(); ++i)
    if (tab[i].bytes != items[i].bytes) PK_FAIL(m, L3_EINVAL, "'%s': tensor %zu has %llu bytes, want %zu", path, i,
                                                (unsigned long long)tab[i].bytes, items[i].bytes);
  // two pinned chunks: the copy of chunk c to the device overlaps the read of chunk c + 1
  Pinned pin[2];
  cudaEvent_t ev[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    PK_CUDA(m, cudaMallocHost(&pin[i].p, kChunk));
    PK_CUDA(m, cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  int cur = 0;
  rc = L3_OK;
  for (size_t i = 0; i < items.size() && rc == L3_OK; ++i) {
    if (fseek(f.f, (long)tab[i].offset, SEEK_SET) != 0) { set_err(m, "'%s': seek failed", path); rc = L3_EINVAL; break; }
    uint64_t cs = 0x243F6A8885A308D3ull;
    for (size_t o = 0; o < items[i].bytes; o += kChunk, cur ^= 1) {
      const size_t n = std::min(kChunk, items[i].bytes - o);
      cudaEventSynchronize(ev[cur]);  // the previous copy out of this chunk has finished
      if (fread(pin[cur].p, 1, n, f.f) != n) { set_err(m, "'%s': truncated at tensor %zu", path, i); rc = L3_EINVAL; break; }
      cs = checksum_update(cs, pin[cur].p, n);
      if (cudaMemcpyAsync((char*)items[i].ptr + o, pin[cur].p, n, cudaMemcpyHostToDevice, m->stream) != cudaSuccess ||
          cudaEventRecord(ev[cur], m->stream) != cudaSuccess) {
        set_err(m, "l3_load_packed: copy to the device failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = L3_ECUDA;
        break;
      }
    }
    if (rc == L3_OK && cs != tab[i].checksum) { set_err(m, "'%s': checksum mismatch in tensor %zu", path, i); rc = L3_EINVAL; }
  }
  cudaStreamSynchronize(m->stream);
  for (int i = 0; i < 2; ++i) cudaEventDestroy(ev[i]);
  if (rc != L3_OK) return rc;  // slots stay "not loaded": l3_finalize refuses a half-filled model
  std::fill(m->loaded.begin(), m->loaded.end(), 1);
  return L3_OK;
}
