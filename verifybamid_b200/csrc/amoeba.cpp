// amoeba.cpp -- see amoeba.h.  Behavioural restatement of the reference's AmoebaMinimizer
// (MathGenMin.cpp:313-443); the element-wise vector arithmetic follows statgen/MathVector.cpp:123-176
// (Add, AddMultiple, Subtract, SetMultiple, Multiply) so that, fed the same function values, the
// simplex visits the same points.
#include "amoeba.h"

#include <cmath>
#include <cstdio>
#include <limits>

namespace vb2 {

namespace {
constexpr double kZeps = 3.0e-10;   // ZEPS,  statgen/MathConstant.h:34
constexpr double kFpMax = 1.0e+100; // FPMAX, statgen/MathConstant.h:36
}  // namespace

void AmoebaMinimizer::Reset(int ndim, double scale) {
  // GeneralMinimizer::Reset (MathGenMin.cpp:17-25): directions = scale * identity, fmin = FPMAX
  directions.assign(ndim, std::vector<double>(ndim, 0.0));
  for (int i = 0; i < ndim; ++i) directions[i][i] = scale;
  point.assign(ndim, 0.0);
  fmin = kFpMax;
  // AmoebaMinimizer::Reset (MathGenMin.cpp:316-324)
  simplex.assign(ndim + 1, std::vector<double>(ndim, 0.0));
  y.assign(ndim + 1, 0.0);
  psum.assign(ndim, 0.0);
  ptry.assign(ndim, 0.0);
}

double AmoebaMinimizer::Minimize(double ftol) {
  const int dim = (int)point.size();
  const int nvertex = dim + 1;
  int ilo, ihi, inhi;
  if (dim == 0) return fmin = f(point);

  // the initial simplex: start + each direction, then the start itself (MathGenMin.cpp:335-345)
  for (int i = 0; i < dim; ++i) {
    simplex[i] = point;
    for (int j = 0; j < dim; ++j) simplex[i][j] += directions[i][j];
    y[i] = f(simplex[i]);
    if (y[i] < fmin) fmin = y[i];
  }
  simplex[nvertex - 1] = point;
  y[nvertex - 1] = f(simplex[nvertex - 1]);
  if (y[nvertex - 1] < fmin) fmin = y[nvertex - 1];
  cycleCount = nvertex;

  auto recompute_psum = [&]() {
    psum = simplex[0];
    for (int m = 1; m < nvertex; ++m)
      for (int j = 0; j < dim; ++j) psum[j] += simplex[m][j];
  };
  recompute_psum();

  while (true) {
    // highest, next-highest and lowest vertex (MathGenMin.cpp:357-370)
    if (y[0] > y[1]) { ilo = inhi = 1; ihi = 0; } else { ilo = inhi = 0; ihi = 1; }
    for (int i = 2; i < nvertex; ++i) {
      if (y[i] <= y[ilo]) ilo = i;
      else if (y[i] > y[ihi]) { inhi = ihi; ihi = i; }
      else if (y[i] > y[inhi]) inhi = i;
    }
    // relative spread of the simplex (MathGenMin.cpp:373-378)
    const double rtol = 2 * std::fabs(y[ihi] - y[ilo]) / (std::fabs(y[ihi]) + std::fabs(y[ilo]) + kZeps);
    if (rtol < ftol) {
      point = simplex[ilo];
      return fmin = y[ilo];
    }
    if (cycleCount > cycleMax) {  // MathGenMin.cpp:380-383
      fprintf(stderr, "WARNING - Amoeba.Minimize - Couldn't converge in %ld cycles\n", cycleMax);
      return std::numeric_limits<double>::max();
    }
    cycleCount += 2;
    double ytry = Amoeba(ihi, -1.0);  // reflect
    if (ytry <= y[ilo]) {
      Amoeba(ihi, 2.0);  // expand
    } else if (ytry >= y[inhi]) {
      const double ysave = y[ihi];
      ytry = Amoeba(ihi, 0.5);  // contract
      if (ytry >= ysave) {      // shrink everything towards the best vertex (MathGenMin.cpp:404-416)
        for (int i = 0; i < nvertex; ++i)
          if (i != ilo) {
            for (int j = 0; j < dim; ++j) simplex[i][j] += simplex[ilo][j];
            for (int j = 0; j < dim; ++j) simplex[i][j] *= 0.5;
            y[i] = f(simplex[i]);
          }
        cycleCount += dim;
        recompute_psum();
      }
    } else {
      cycleCount--;
    }
  }
}

double AmoebaMinimizer::Amoeba(int ihi, double factor) {
  const int dim = (int)point.size();
  const double fac = (1.0 - factor) / dim;
  for (int i = 0; i < dim; ++i) ptry[i] = fac * psum[i];                      // SetMultiple
  for (int i = 0; i < dim; ++i) ptry[i] += (factor - fac) * simplex[ihi][i];  // AddMultiple
  const double ytry = f(ptry);
  if (ytry < y[ihi]) {
    y[ihi] = ytry;
    for (int i = 0; i < dim; ++i) psum[i] -= simplex[ihi][i];
    simplex[ihi] = ptry;
    for (int i = 0; i < dim; ++i) psum[i] += simplex[ihi][i];
  }
  return ytry;
}

}  // namespace vb2
