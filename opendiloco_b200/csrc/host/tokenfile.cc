// Pre-tokenised corpus reader (C ABI): memory-maps token shards and serves fixed-length causal-LM batches through the
// prefetch ring, so the training loop's per-micro-batch host work stays one async H2D copy at any token rate.
//
// Why it exists: the reference tokenises C4 on the fly inside 4 DataLoader workers (train_fsdp.py:132-168).  That feeds a
// few hundred thousand tokens per second; this framework consumes ~0.78 M tokens/s per GPU (6 M/s per box), so the
// production input path is "tokenise once (scripts/tokenize_corpus.py), stream the shards".
//
// Shard format: 32-byte header { char magic[8] = "ODBTOK1\0"; u32 bytes_per_token (2|4); u32 reserved; u64 n_tokens;
// u64 reserved } followed by the tokens (little endian).  Files without the magic are taken as raw token arrays of
// `raw_bytes_per_token` (nanoGPT-style .bin).
//
// Sampling: the corpus is cut into non-overlapping windows of seq_len tokens (per shard; a shard's tail is dropped).
// Sample g of the global stream = window perm_e(g mod W) of epoch e = g / W, where perm_e is a seeded bijection
// (4-round Feistel network with cycle walking) or the identity.  Batch b of rank r holds samples (b*world + r)*B + i:
// ranks read disjoint samples, every window is visited once per epoch, and the position (one integer) is the whole state.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "prefetch_ring.h"

#define ODB_API extern "C" __attribute__((visibility("default")))

using odbhost::PrefetchRing;
using odbhost::splitmix64;

namespace {

struct Shard {
  const uint8_t* base = nullptr;      // first token
  void* map = nullptr;
  size_t map_len = 0;
  int bytes = 2;
  int64_t n_tokens = 0, n_windows = 0, first_window = 0;
};

struct Corpus {
  std::vector<Shard> shards;
  int64_t windows = 0, seq = 0, batch = 0, rank = 0, world = 1;
  uint64_t seed = 0;
  bool shuffle = true;
  int bits = 1;                       // Feistel domain: 2^bits >= windows, bits even

  ~Corpus() {
    for (auto& s : shards)
      if (s.map) ::munmap(s.map, s.map_len);
  }

  // seeded bijection on [0, windows): balanced Feistel over `bits` bits, cycle-walked into range
  int64_t permute(int64_t x, int64_t epoch) const {
    if (!shuffle || windows <= 1) return x;
    const int half = bits / 2;
    const uint64_t mask = (1ull << half) - 1;
    uint64_t v = (uint64_t)x;
    do {
      uint64_t l = v >> half, r = v & mask;
      for (int round = 0; round < 4; ++round) {
        const uint64_t f = splitmix64(r ^ splitmix64(seed + 0x51ED270B9A7F3C21ull * (uint64_t)(epoch + 1) + (uint64_t)round)) & mask;
        const uint64_t nl = r, nr = l ^ f;
        l = nl;
        r = nr;
      }
      v = (l << half) | r;
    } while (v >= (uint64_t)windows);
    return (int64_t)v;
  }

  void read_window(int64_t w, int64_t* dst) const {
    size_t lo = 0, hi = shards.size() - 1;                 // last shard whose first_window <= w
    while (lo < hi) {
      const size_t mid = (lo + hi + 1) / 2;
      if (shards[mid].first_window <= w) lo = mid; else hi = mid - 1;
    }
    const Shard& s = shards[lo];
    const int64_t off = (w - s.first_window) * seq;
    if (s.bytes == 2) {
      const uint16_t* p = reinterpret_cast<const uint16_t*>(s.base) + off;
      for (int64_t i = 0; i < seq; ++i) dst[i] = p[i];
    } else {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(s.base) + off;
      for (int64_t i = 0; i < seq; ++i) dst[i] = p[i];
    }
  }

  void fill(int64_t* dst, int64_t b) const {
    for (int64_t i = 0; i < batch; ++i) {
      const int64_t g = (b * world + rank) * batch + i;
      read_window(permute(g % windows, g / windows), dst + i * seq);
    }
  }
};

bool open_shard(const char* path, int raw_bytes, int64_t seq, Shard* out) {
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (::fstat(fd, &st) != 0 || st.st_size <= 0) { ::close(fd); return false; }
  void* m = ::mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
  ::close(fd);
  if (m == MAP_FAILED) return false;
  ::madvise(m, (size_t)st.st_size, MADV_RANDOM);
  const uint8_t* p = static_cast<const uint8_t*>(m);
  out->map = m;
  out->map_len = (size_t)st.st_size;
  if (st.st_size >= 32 && memcmp(p, "ODBTOK1", 8) == 0) {
    uint32_t bytes;
    uint64_t n;
    memcpy(&bytes, p + 8, 4);
    memcpy(&n, p + 16, 8);
    if ((bytes != 2 && bytes != 4) || 32 + n * bytes > (uint64_t)st.st_size) { ::munmap(m, out->map_len); out->map = nullptr; return false; }
    out->bytes = (int)bytes;
    out->n_tokens = (int64_t)n;
    out->base = p + 32;
  } else {
    if (raw_bytes != 2 && raw_bytes != 4) { ::munmap(m, out->map_len); out->map = nullptr; return false; }
    out->bytes = raw_bytes;
    out->n_tokens = (int64_t)(st.st_size / raw_bytes);
    out->base = p;
  }
  out->n_windows = out->n_tokens / seq;
  return true;
}

}  // namespace

// paths: `n_paths` NUL-terminated shard paths.  Returns a PrefetchRing handle served by odb_tg_next / _release / _position /
// _destroy, or nullptr (bad file, or fewer windows than one global batch).  *windows_out receives the windows per epoch.
ODB_API void* odb_tf_open(const char** paths, int n_paths, int raw_bytes_per_token, int64_t seq_len, int64_t batch, int64_t rank,
                          int64_t world, uint64_t seed, int shuffle, int nbuf, int64_t** bufs, int64_t start_batch,
                          int64_t* windows_out) {
  auto* c = new Corpus();
  c->seq = seq_len; c->batch = batch; c->rank = rank; c->world = world; c->seed = seed; c->shuffle = shuffle != 0;
  for (int i = 0; i < n_paths; ++i) {
    Shard s;
    if (!open_shard(paths[i], raw_bytes_per_token, seq_len, &s)) { delete c; return nullptr; }
    s.first_window = c->windows;
    c->windows += s.n_windows;
    if (s.n_windows > 0) c->shards.push_back(s);
    else if (s.map) ::munmap(s.map, s.map_len);
  }
  if (windows_out) *windows_out = c->windows;
  if (c->windows < batch * world || c->shards.empty()) { delete c; return nullptr; }
  c->bits = 2;
  while ((1ll << c->bits) < c->windows) c->bits += 2;
  auto* ring = new PrefetchRing();
  ring->fill = [c](int64_t* dst, int64_t b) { c->fill(dst, b); };
  ring->on_destroy = [c] { delete c; };
  ring->start(bufs, nbuf, start_batch);
  return ring;
}

// Writes one shard (header + tokens); bytes_per_token 2 or 4.  Returns 0 on success.
ODB_API int odb_tf_write(const char* path, const int64_t* tokens, int64_t n, int bytes_per_token) {
  if (bytes_per_token != 2 && bytes_per_token != 4) return -1;
  const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return -2;
  uint8_t hdr[32] = {0};
  memcpy(hdr, "ODBTOK1", 8);
  const uint32_t b = (uint32_t)bytes_per_token;
  const uint64_t cnt = (uint64_t)n;
  memcpy(hdr + 8, &b, 4);
  memcpy(hdr + 16, &cnt, 8);
  bool ok = ::write(fd, hdr, 32) == 32;
  std::vector<uint8_t> chunk;
  const int64_t step = 1 << 20;
  for (int64_t i = 0; ok && i < n; i += step) {
    const int64_t m = (n - i) < step ? (n - i) : step;
    chunk.resize((size_t)m * bytes_per_token);
    for (int64_t j = 0; j < m; ++j) {
      const int64_t t = tokens[i + j];
      if (t < 0 || (bytes_per_token == 2 && t > 0xFFFF) || t > 0xFFFFFFFFll) { ok = false; break; }
      if (bytes_per_token == 2) { const uint16_t v = (uint16_t)t; memcpy(&chunk[(size_t)j * 2], &v, 2); }
      else { const uint32_t v = (uint32_t)t; memcpy(&chunk[(size_t)j * 4], &v, 4); }
    }
    if (ok) ok = ::write(fd, chunk.data(), chunk.size()) == (ssize_t)chunk.size();
  }
  ::close(fd);
  if (!ok) ::unlink(path);             // never leave a truncated shard behind
  return ok ? 0 : -3;
}
