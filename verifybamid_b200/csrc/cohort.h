// cohort.h -- lock-step likelihood evaluation for a cohort of samples (BASELINE.json configs[3]: "batch of
// 64 synthetic samples, one sample per block").
//
// Every sample runs the reference's own sequential optimisation (OptimizeLLK -> AmoebaMinimizer) on its own
// host thread, unchanged.  What changes is who launches: a sample thread that needs a likelihood hands
// (context, parameters) to the coordinator and sleeps; once every still-running sample of the device has
// asked, the coordinator evaluates all of them with ONE launch (vb2_llk_eval_many, one job per sample) and
// wakes them.  Each simplex therefore follows exactly the trajectory it would follow alone (the many-sample
// launch returns the same bits as a single evaluation), while the GPU sees one wide launch per step instead
// of one narrow launch per sample and step.  Samples are independent: there is no collective, and one
// coordinator per GPU scales the cohort across devices.
#ifndef VB2_COHORT_H_
#define VB2_COHORT_H_

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "vb2_llk.h"

namespace vb2 {

class CohortCoordinator {
 public:
  // One launch: (contexts, n, pc_contam[n][k], pc_intended[n][k], alphas[n], out[n]) -> VB2 status.
  // The default is vb2_llk_eval_many; tests of the lock-step logic substitute a host function.
  typedef std::function<int(vb2_llk_ctx *const *, int, const double *, const double *, const double *, double *)> Launcher;
  CohortCoordinator(int n_samples, int n_pc, Launcher launcher = Launcher());
  // sample threads ------------------------------------------------------------------------------
  double Evaluate(int sample, vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha);
  void Finish(int sample);  // the sample's optimisation is over (also on error)
  // coordinator thread --------------------------------------------------------------------------
  void Run();               // returns when every sample has finished
  long launches = 0, evaluations = 0;
  std::string error;        // first engine error, if any (then every Evaluate throws)

 private:
  enum State { kIdle, kWaiting, kDone };
  const int n_, k_;
  Launcher launch_;
  std::mutex mu_;
  std::condition_variable cv_request_, cv_result_;
  std::vector<State> state_;
  std::vector<vb2_llk_ctx *> ctx_;
  std::vector<double> pc1_, pc2_, alpha_, result_;
  std::vector<unsigned long> serial_;  // bumped when result_[i] is filled
  int waiting_ = 0, done_ = 0;
};

}  // namespace vb2
#endif
