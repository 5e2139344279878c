// cohort.cpp -- see cohort.h.
#include "cohort.h"

#include <algorithm>
#include <stdexcept>

namespace vb2 {

CohortCoordinator::CohortCoordinator(int n_samples, int n_pc, Launcher launcher)
    : n_(n_samples), k_(n_pc), launch_(launcher ? launcher : Launcher(vb2_llk_eval_many)), state_(n_samples, kIdle), ctx_(n_samples, nullptr), pc1_((size_t)n_samples * n_pc),
      pc2_((size_t)n_samples * n_pc), alpha_(n_samples), result_(n_samples), serial_(n_samples, 0) {}

double CohortCoordinator::Evaluate(int i, vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended,
                                   double alpha) {
  std::unique_lock<std::mutex> lock(mu_);
  if (!error.empty()) throw std::runtime_error(error);
  ctx_[i] = ctx;
  for (int d = 0; d < k_; ++d) {
    pc1_[(size_t)i * k_ + d] = pc_contam[d];
    pc2_[(size_t)i * k_ + d] = pc_intended[d];
  }
  alpha_[i] = alpha;
  state_[i] = kWaiting;
  ++waiting_;
  const unsigned long ticket = serial_[i];
  cv_request_.notify_one();
  cv_result_.wait(lock, [&] { return serial_[i] != ticket || !error.empty(); });
  if (serial_[i] == ticket) throw std::runtime_error(error);
  return result_[i];
}

void CohortCoordinator::Finish(int i) {
  std::lock_guard<std::mutex> lock(mu_);
  if (state_[i] == kDone) return;
  if (state_[i] == kWaiting) --waiting_;
  state_[i] = kDone;
  ++done_;
  cv_request_.notify_one();
}

void CohortCoordinator::Run() {
  std::vector<int> who;
  std::vector<vb2_llk_ctx *> ctxs;
  std::vector<double> pc1, pc2, al, out;
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    // a step is complete when every sample that is still running has asked for its next likelihood
    cv_request_.wait(lock, [&] { return done_ == n_ || (waiting_ > 0 && waiting_ + done_ == n_); });
    if (done_ == n_) return;
    who.clear(); ctxs.clear(); pc1.clear(); pc2.clear(); al.clear();
    for (int i = 0; i < n_; ++i)
      if (state_[i] == kWaiting) {
        who.push_back(i);
        ctxs.push_back(ctx_[i]);
        pc1.insert(pc1.end(), pc1_.begin() + (size_t)i * k_, pc1_.begin() + (size_t)(i + 1) * k_);
        pc2.insert(pc2.end(), pc2_.begin() + (size_t)i * k_, pc2_.begin() + (size_t)(i + 1) * k_);
        al.push_back(alpha_[i]);
      }
    out.assign(who.size(), 0.0);
    lock.unlock();
    // one launch: job j evaluates sample who[j] (processed in VB2_MAX_BATCH pieces for very large cohorts)
    int rc = VB2_OK;
    for (size_t b = 0; b < who.size() && rc == VB2_OK; b += VB2_MAX_BATCH) {
      const int m = (int)std::min<size_t>(VB2_MAX_BATCH, who.size() - b);
      rc = launch_(ctxs.data() + b, m, pc1.data() + b * k_, pc2.data() + b * k_, al.data() + b, out.data() + b);
      ++launches;
    }
    lock.lock();
    if (rc != VB2_OK) {
      error = std::string("GPU engine (cohort launch): ") + vb2_last_error(ctxs.empty() ? nullptr : ctxs[0]);
      cv_result_.notify_all();
      // keep serving Finish() calls: every sample thread will now throw out of Evaluate and finish
      cv_request_.wait(lock, [&] { return done_ == n_; });
      return;
    }
    evaluations += (long)who.size();
    for (size_t j = 0; j < who.size(); ++j) {
      const int i = who[j];
      result_[i] = out[j];
      ++serial_[i];
      state_[i] = kIdle;
      --waiting_;
    }
    cv_result_.notify_all();
  }
}

}  // namespace vb2
