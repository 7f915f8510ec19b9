// Batching rounds: a barrier for host threads that each repeatedly produce one request and need it executed together with the
// requests of all other active threads (QJMC ensembles: every worker's next truncated SVD; tn_svd.cu).
//   submit(req, exec): parks the request; when every active participant is parked, the last arrival calls
//                      exec(std::vector<Req*>&) once (with the lock held, so exec never runs concurrently) and wakes the others.
//   leave(exec):       the caller stops participating; completes a round if everybody else is already parked.
// An exception thrown by exec is remembered and re-thrown (as std::runtime_error with the same text) by every submit of that and
// all later rounds, so no participant can be left waiting.
// Plain C++ (no CUDA) so that the synchronisation logic is unit-tested on the CPU: tests/cpu/test_rounds.cpp.
#pragma once
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace tn {

template <class Req>
class Rounds {
 public:
  explicit Rounds(int nactive) : nactive_(nactive) {}

  template <class Exec>
  void submit(Req* r, Exec&& exec) {
    std::unique_lock<std::mutex> lk(mu_);
    if (failed_) throw std::runtime_error(err_);
    pending_.push_back(r);
    const unsigned long long my = round_;
    if ((int)pending_.size() >= nactive_) run(exec);
    else cv_.wait(lk, [&] { return round_ != my; });
    if (failed_) throw std::runtime_error(err_);
  }

  template <class Exec>
  void leave(Exec&& exec) {
    std::unique_lock<std::mutex> lk(mu_);
    --nactive_;
    if (nactive_ > 0 && !pending_.empty() && (int)pending_.size() >= nactive_) run(exec);
  }

  unsigned long long rounds() const { return round_; }
  long long requests() const { return requests_; }
  bool failed() const { return failed_; }

 private:
  template <class Exec>
  void run(Exec&& exec) {      // lock held
    if (!failed_) {
      try { exec(pending_); }
      catch (const std::exception& e) { failed_ = true; err_ = e.what(); }
      catch (...) { failed_ = true; err_ = "unknown error in a batching round"; }
    }
    requests_ += (long long)pending_.size();
    pending_.clear();
    ++round_;
    cv_.notify_all();
  }

  std::mutex mu_;
  std::condition_variable cv_;
  int nactive_;
  unsigned long long round_ = 0;
  long long requests_ = 0;
  std::vector<Req*> pending_;
  bool failed_ = false;
  std::string err_;
};

}  // namespace tn
