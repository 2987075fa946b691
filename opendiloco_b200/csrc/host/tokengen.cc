// Synthetic-token prefetcher (C ABI): a background thread keeps a ring of caller-provided (pinned) int64 buffers
// filled with uniform tokens in [lo, hi), so the training loop's per-micro-batch host work is one async H2D copy.
// Token law = the reference's FakeTokenizedDataset (utils.py:155-167).  The stream is a pure function of
// (seed, batch index): counter-based generation makes checkpoint/resume exact (state = next batch index).
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#define ODB_API extern "C" __attribute__((visibility("default")))

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Gen {
  uint64_t seed;
  int64_t lo, hi, elems;
  std::vector<int64_t*> bufs;
  std::vector<int64_t> batch_of;      // batch index held by each buffer, -1 = free
  int64_t next_fill = 0, next_take = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::thread worker;
  bool stop = false;

  void fill(int64_t* dst, int64_t batch) const {
    const uint64_t range = (uint64_t)(hi - lo);
    const uint64_t base = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(batch + 1)));
    for (int64_t i = 0; i < elems; ++i) {
      const uint64_t r = splitmix64(base + (uint64_t)i);
      dst[i] = lo + (int64_t)(((unsigned __int128)r * range) >> 64);   // unbiased enough for range << 2^64
    }
  }
  void run() {
    std::unique_lock<std::mutex> lk(mu);
    while (!stop) {
      int slot = -1;
      for (size_t s = 0; s < bufs.size(); ++s)
        if (batch_of[s] < 0) { slot = (int)s; break; }
      if (slot < 0) { cv.wait(lk); continue; }
      const int64_t b = next_fill++;
      batch_of[slot] = -2;                 // being filled
      lk.unlock();
      fill(bufs[slot], b);
      lk.lock();
      batch_of[slot] = b;
      cv.notify_all();
    }
  }
};

}  // namespace

ODB_API void* odb_tg_create(uint64_t seed, int64_t lo, int64_t hi, int64_t elems, int nbuf, int64_t** bufs, int64_t start_batch) {
  Gen* g = new Gen();
  g->seed = seed; g->lo = lo; g->hi = hi; g->elems = elems;
  g->bufs.assign(bufs, bufs + nbuf);
  g->batch_of.assign(nbuf, -1);
  g->next_fill = g->next_take = start_batch;
  g->worker = std::thread([g] { g->run(); });
  return g;
}
// Blocks until the next batch (in order) is ready; returns its buffer slot.
ODB_API int odb_tg_next(void* h) {
  Gen* g = (Gen*)h;
  std::unique_lock<std::mutex> lk(g->mu);
  const int64_t want = g->next_take;
  for (;;) {
    for (size_t s = 0; s < g->bufs.size(); ++s)
      if (g->batch_of[s] == want) { g->next_take++; return (int)s; }
    g->cv.wait(lk);
  }
}
ODB_API void odb_tg_release(void* h, int slot) {
  Gen* g = (Gen*)h;
  std::lock_guard<std::mutex> lk(g->mu);
  g->batch_of[slot] = -1;
  g->cv.notify_all();
}
ODB_API int64_t odb_tg_position(void* h) { Gen* g = (Gen*)h; std::lock_guard<std::mutex> lk(g->mu); return g->next_take; }
// Direct (synchronous) generation of one batch - used by tests and by the Python fallback for parity.
ODB_API void odb_tg_fill(uint64_t seed, int64_t lo, int64_t hi, int64_t elems, int64_t batch, int64_t* dst) {
  Gen g; g.seed = seed; g.lo = lo; g.hi = hi; g.elems = elems;
  g.fill(dst, batch);
}
ODB_API void odb_tg_destroy(void* h) {
  Gen* g = (Gen*)h;
  { std::lock_guard<std::mutex> lk(g->mu); g->stop = true; }
  g->cv.notify_all();
  if (g->worker.joinable()) g->worker.join();
  delete g;
}
