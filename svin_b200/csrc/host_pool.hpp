// Persistent host worker threads shared by the engines' upload / download paths.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace svin {

// Persistent host worker threads of one context: the upload path (ordering, packing) and the result scatter are
// independent per window / problem, and spawning threads on every call costs more than the work itself.
class HostPool {
 public:
  explicit HostPool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this] { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }
  int workers() const { return (int)th_.size(); }
  // hand items 0..n-1 to the workers; the caller may help() and must wait()
  void start(int n, std::function<void(int)> fn) {
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = std::move(fn);
      n_ = n;
      next_.store(0);
      running_ = (int)th_.size();
      ++gen_;
    }
    cv_work_.notify_all();
  }
  void help() {
    for (int i = next_.fetch_add(1); i < n_; i = next_.fetch_add(1)) fn_(i);
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return running_ == 0; });
  }
  void run(int n, std::function<void(int)> fn) {
    start(n, std::move(fn));
    help();
    wait();
  }

 private:
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m_);
      cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      lk.unlock();
      help();
      lk.lock();
      if (--running_ == 0) cv_done_.notify_all();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  std::function<void(int)> fn_;
  int n_ = 0;
  std::atomic<int> next_{0};
  int running_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};


}  // namespace svin
