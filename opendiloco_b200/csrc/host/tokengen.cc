// Synthetic-token prefetcher (C ABI): a background thread keeps a ring of caller-provided (pinned) int64 buffers
// filled with uniform tokens in [lo, hi), so the training loop's per-micro-batch host work is one async H2D copy.
// Token law = the reference's FakeTokenizedDataset (utils.py:155-167).  The stream is a pure function of
// (seed, batch index): counter-based generation makes checkpoint/resume exact (state = next batch index).
#include <stdint.h>

#include "prefetch_ring.h"

#define ODB_API extern "C" __attribute__((visibility("default")))

using odbhost::PrefetchRing;
using odbhost::splitmix64;

namespace {

void fill_uniform(uint64_t seed, int64_t lo, int64_t hi, int64_t elems, int64_t* dst, int64_t batch) {
  const uint64_t range = (uint64_t)(hi - lo);
  const uint64_t base = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(batch + 1)));
  for (int64_t i = 0; i < elems; ++i) {
    const uint64_t r = splitmix64(base + (uint64_t)i);
    dst[i] = lo + (int64_t)(((unsigned __int128)r * range) >> 64);   // unbiased enough for range << 2^64
  }
}

}  // namespace

// The handle returned by odb_tg_create (and by odb_tf_open in tokenfile.cc) is a PrefetchRing: next / release / position /
// destroy below serve both sources.
ODB_API void* odb_tg_create(uint64_t seed, int64_t lo, int64_t hi, int64_t elems, int nbuf, int64_t** bufs, int64_t start_batch) {
  auto* g = new PrefetchRing();
  g->fill = [=](int64_t* dst, int64_t batch) { fill_uniform(seed, lo, hi, elems, dst, batch); };
  g->start(bufs, nbuf, start_batch);
  return g;
}
// Blocks until the next batch (in order) is ready; returns its buffer slot.
ODB_API int odb_tg_next(void* h) { return static_cast<PrefetchRing*>(h)->next(); }
ODB_API void odb_tg_release(void* h, int slot) { static_cast<PrefetchRing*>(h)->release(slot); }
ODB_API int64_t odb_tg_position(void* h) { return static_cast<PrefetchRing*>(h)->position(); }
// Direct (synchronous) generation of one batch - used by tests and by the Python fallback for parity.
ODB_API void odb_tg_fill(uint64_t seed, int64_t lo, int64_t hi, int64_t elems, int64_t batch, int64_t* dst) {
  fill_uniform(seed, lo, hi, elems, dst, batch);
}
ODB_API void odb_tg_destroy(void* h) { delete static_cast<PrefetchRing*>(h); }
