// kaldi-hmm-gmm_b200/csrc/khg_host_pool.h — the library's small persistent pool of host threads: the aligner's passes
// over the graphs (khg_align.cu) and the staging of large pageable host buffers into pinned memory (khg_b200.cu).
#pragma once
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace khg {

// A small persistent pool for the host passes over the graphs (a call makes ~10 of them; creating 16 threads each
// time costs more than the passes themselves at C5 sizes).  One job at a time (callers are serialised by a mutex).
class HostPool {
 public:
  static HostPool &get() {
    static HostPool p;
    return p;
  }
  int workers() const { return (int)th_.size() + 1; }
  // runs job(w) for w = 0 .. n_workers-1 (the caller is worker 0) and returns when all are done
  void run(int n_workers, const std::function<void(int)> &job) {
    std::lock_guard<std::mutex> one(call_mu_);
    n_workers = std::max(1, std::min(n_workers, workers()));
    if (getpid() != pid_) n_workers = 1;  // (a forked child has no worker threads: the caller does everything)
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &job;
      active_ = n_workers - 1;
      pending_ = n_workers - 1;
      ++epoch_;
    }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  HostPool() {
    const int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    for (int w = 1; w < nt; ++w) th_.emplace_back([this, w] { loop(w); });
    pid_ = getpid();
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  void loop(int w) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)> *job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        if (w <= active_) job = job_;
      }
      if (job) {
        (*job)(w);
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)> *job_ = nullptr;
  int active_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
  pid_t pid_ = 0;
};

template <class F>
inline void parallel_for(int n, F f, int min_parallel = 64) {
  int nt = std::min(HostPool::get().workers(), std::max(1, n));
  if (n < min_parallel) nt = 1;
  if (nt == 1) {
    for (int i = 0; i < n; ++i) f(i, 0);
    return;
  }
  // (items are handed out dynamically: in a forked child the caller alone runs, and takes them all)
  std::atomic<int> next{0};
  HostPool::get().run(nt, [&](int w) {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i, w);
  });
}

// memcpy of a large block by several of the pool's threads (1 MB pieces); small blocks are a plain memcpy
inline void *parallel_memcpy(void *dst, const void *src, size_t bytes, int max_threads = 8) {
  constexpr size_t kPiece = 1u << 20;
  const int threads = std::min(max_threads, HostPool::get().workers());
  if (bytes < 8 * kPiece || threads <= 1) return std::memcpy(dst, src, bytes);
  const int n_pieces = (int)((bytes + kPiece - 1) / kPiece);
  std::atomic<int> next{0};
  HostPool::get().run(threads, [&](int) {
    for (int q = next.fetch_add(1); q < n_pieces; q = next.fetch_add(1))
      std::memcpy(static_cast<char *>(dst) + (size_t)q * kPiece, static_cast<const char *>(src) + (size_t)q * kPiece,
                  std::min(kPiece, bytes - (size_t)q * kPiece));
  });
  return dst;
}

}  // namespace khg
