// Ring of caller-provided (pinned) int64 batch buffers kept full by one background thread.  The batch source is any
// pure function (dst, batch index) -> tokens, so the consumer's position - one integer - is the whole resumable state.
// Shared by the synthetic generator (tokengen.cc) and the pre-tokenised corpus reader (tokenfile.cc).
#pragma once
#include <stdint.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace odbhost {

struct PrefetchRing {
  std::function<void(int64_t*, int64_t)> fill;   // (destination buffer, batch index)
  std::function<void()> on_destroy;              // releases whatever the source owns (mappings, ...)
  std::vector<int64_t*> bufs;
  std::vector<int64_t> batch_of;                 // batch index held by each buffer; -1 = free, -2 = being filled
  int64_t next_fill = 0, next_take = 0;
  std::mutex mu;
  std::condition_variable cv;
  std::thread worker;
  bool stop = false;

  void start(int64_t** b, int nbuf, int64_t start_batch) {
    bufs.assign(b, b + nbuf);
    batch_of.assign(nbuf, -1);
    next_fill = next_take = start_batch;
    worker = std::thread([this] { run(); });
  }

  void run() {
    std::unique_lock<std::mutex> lk(mu);
    while (!stop) {
      int slot = -1;
      for (size_t s = 0; s < bufs.size(); ++s)
        if (batch_of[s] == -1) { slot = (int)s; break; }
      if (slot < 0) { cv.wait(lk); continue; }
      const int64_t b = next_fill++;
      batch_of[slot] = -2;
      lk.unlock();
      fill(bufs[slot], b);
      lk.lock();
      batch_of[slot] = b;
      cv.notify_all();
    }
  }

  // blocks until the next batch (in order) is ready; returns its slot
  int next() {
    std::unique_lock<std::mutex> lk(mu);
    const int64_t want = next_take;
    for (;;) {
      for (size_t s = 0; s < bufs.size(); ++s)
        if (batch_of[s] == want) { next_take++; return (int)s; }
      cv.wait(lk);
    }
  }

  void release(int slot) {
    std::lock_guard<std::mutex> lk(mu);
    batch_of[slot] = -1;
    cv.notify_all();
  }

  int64_t position() {
    std::lock_guard<std::mutex> lk(mu);
    return next_take;
  }

  ~PrefetchRing() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    if (worker.joinable()) worker.join();
    if (on_destroy) on_destroy();
  }
};

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace odbhost
