// amoeba.h -- Nelder-Mead downhill simplex with the exact control flow of the reference's
// AmoebaMinimizer (MathGenMin.cpp:313-443, MathGenMin.h:92-108) and GeneralMinimizer::Reset
// (MathGenMin.cpp:17-25): the caller of the likelihood hot path.  Host code; it decides how many
// evaluations a run needs, so its branching must not drift from the reference.
#ifndef VB2_AMOEBA_H_
#define VB2_AMOEBA_H_

#include <vector>

namespace vb2 {

// The seam of the reference: VectorFunc::Evaluate(Vector&) -> double (statgen/MathVector.h:281-308).
class VectorFunc {
 public:
  virtual ~VectorFunc() {}
  virtual double Evaluate(const std::vector<double> &v) = 0;
};

class AmoebaMinimizer {
 public:
  VectorFunc *func = nullptr;                   // GeneralMinimizer::func
  std::vector<double> point;                    // start on entry, best vertex on convergence
  double fmin = 1.0e+100;                       // FPMAX, statgen/MathConstant.h:36
  long cycleCount = 0, cycleMax = 50000;        // MathGenMin.cpp:313-314
  std::vector<std::vector<double>> simplex;     // (dim+1) x dim
  std::vector<std::vector<double>> directions;  // dim x dim, identity * scale after Reset

  void Reset(int ndim, double scale = 1.0);     // MathGenMin.cpp:316-324
  // Returns the minimum found, or std::numeric_limits<double>::max() when cycleMax is exceeded.
  double Minimize(double ftol);                 // MathGenMin.cpp:326-423

 private:
  std::vector<double> psum, ptry, y;
  double Amoeba(int ihi, double factor);        // MathGenMin.cpp:425-443
  double f(const std::vector<double> &v) { return func->Evaluate(v); }  // MathGenMin.h:30-31
};

}  // namespace vb2
#endif
