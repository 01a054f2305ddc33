// Persistent host worker pool shared by the host-only routines of the library (result expansion, contours).
#pragma once
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mrcnn_host {

// ---- persistent worker pool -------------------------------------------------------------------
class Pool {
 public:
  static Pool& get() {
    static Pool p;
    return p;
  }
  // runs fn(worker_index) on `n` threads (the caller is worker 0) and returns when all are done
  void run(int n, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mu_);      // one parallel region at a time
    if (n <= 1) {
      fn(0);
      return;
    }
    ensure(n - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      want_ = n - 1;
      pending_ = n - 1;
      ++epoch_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void ensure(int n) {
    while ((int)threads_.size() < n) {
      const int id = (int)threads_.size();
      uint64_t start_epoch;
      {
        std::lock_guard<std::mutex> lk(mu_);
        start_epoch = epoch_;
      }
      threads_.emplace_back([this, id, start_epoch] { worker(id, start_epoch); });
    }
  }
  void worker(int id, uint64_t seen) {
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        if (id < want_) fn = fn_;
      }
      if (fn) {
        (*fn)(id + 1);
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::mutex run_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void(int)>* fn_ = nullptr;
  int want_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

inline int default_threads() {
  static int n = 0;
  if (n == 0) {
    int avail = 1;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) avail = CPU_COUNT(&set);
    if (avail < 1) avail = 1;
    n = avail > 16 ? 16 : avail;
    if (const char* e = getenv("MRCNN_B200_HOST_THREADS")) {
      const int v = atoi(e);
      if (v >= 1 && v <= 256) n = v;
    }
  }
  return n;
}


}  // namespace mrcnn_host
