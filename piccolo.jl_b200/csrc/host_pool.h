// Small persistent host thread pool used by the host-pointer entry points to un-pack compact
// device records into the caller's arrays while the next chunk is still crossing PCIe.
// Pure data movement (replicating already computed values): no arithmetic of the path runs here.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pb2 {

class HostPool {
 public:
  static HostPool& instance() {
    static HostPool pool;
    return pool;
  }
  int threads() const { return (int)workers_.size() + 1; }

  // fn(first, last) over [0, n) in blocks of `grain`, handed out in increasing order.  begin() wakes the
  // workers and returns; the caller may do other things (e.g. wait for DMA chunks the blocks depend
  // on), then finish() makes it take part and returns when every block is done.
  void begin(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)>& fn) {
    call_mutex_.lock();   // one job at a time
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn; n_ = n; grain_ = grain;
      next_.store(0, std::memory_order_relaxed);
      pending_.store((int)workers_.size(), std::memory_order_relaxed);
      ++generation_;
    }
    cv_.notify_all();
  }
  void finish() {
    run();
    while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    call_mutex_.unlock();
  }

 private:
  HostPool() {
    // PB2_HOST_THREADS = threads that take part in un-packing, the caller included (default: all cores, at most
    // 8).  Several processes sharing one host (one rank per GPU) should divide the cores between them: the
    // workers spin on the arrival of DMA chunks, so over-subscription costs more than it gains.
    unsigned hw = std::thread::hardware_concurrency();
    int n = hw > 1 ? (int)std::min<unsigned>(hw - 1, 7) : 0;
    if (const char* env = std::getenv("PB2_HOST_THREADS")) n = std::max(0, std::min(63, std::atoi(env) - 1));
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void run() {
    for (;;) {
      const int64_t b = next_.fetch_add(grain_, std::memory_order_relaxed);
      if (b >= n_) break;
      (*fn_)(b, std::min(b + grain_, n_));
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      run();
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, call_mutex_;
  std::condition_variable cv_;
  uint64_t generation_ = 0;
  bool stop_ = false;
  const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
  int64_t n_ = 0, grain_ = 1;
  std::atomic<int64_t> next_{0};
  std::atomic<int> pending_{0};
};

}  // namespace pb2
