// host_shim.cpp -- C entry points over the HOST side of the engine (file readers, text-pileup parser,
// marker resolution, depth sanity filter, Nelder-Mead) so that it can be exercised without a GPU
// (tests/test_host.py) and reused from Python (verifybamid_b200/host.py).  Built into libvb2host.so.
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include "amoeba.h"
#include "estimator.h"

using namespace vb2;

namespace {
struct CallbackFunc : VectorFunc {
  double (*fn)(void *, const double *, int);
  void *user;
  double Evaluate(const std::vector<double> &v) override { return fn(user, v.data(), (int)v.size()); }
};
struct Loaded {
  ContaminationEstimator *E = nullptr;
  int sanity_ok = 1;
  std::string err;
};
}  // namespace

extern "C" {

// AmoebaMinimizer::Reset(dim) + Minimize(ftol) on an arbitrary callback; `point` is start in, best out.
double vb2_host_amoeba_minimize(double (*fn)(void *, const double *, int), void *user, int dim, double *point,
                                double ftol, long *cycles) {
  CallbackFunc f;
  f.fn = fn;
  f.user = user;
  AmoebaMinimizer m;
  m.func = &f;
  m.Reset(dim);
  m.point.assign(point, point + dim);
  double r = m.Minimize(ftol);
  for (int i = 0; i < dim; ++i) point[i] = m.point[i];
  if (cycles) *cycles = m.cycleCount;
  return r;
}

// The CLI's load sequence up to (and including) the sanity check and BuildResolvedMarkers
// (reference main.cpp:283-333, :371-379; ContaminationEstimator.cpp:67-86).  NULL on failure.
void *vb2_host_load(const char *svd_prefix, const char *pileup, int n_pc, int disable_sanity, const char *known_af) {
  Loaded *L = new Loaded();
  try {
    std::string p(svd_prefix);
    L->E = new ContaminationEstimator(n_pc, (p + ".bed").c_str(), 4, 1e-8);
    L->E->isSanityCheckDisabled = disable_sanity != 0;
    if (known_af && *known_af) {
      L->E->isAFknown = true;
      L->E->isPCFixed = true;
      L->E->isHeter = false;
      L->E->ReadAF(known_af);
    }
    L->E->ReadSVDMatrix(p + ".UD", p + ".V", p + ".mu");
    L->E->ReadPileup(pileup);
    if (!disable_sanity) L->sanity_ok = L->E->IsSanityCheckOK() ? 1 : 0;
    L->E->BuildResolvedMarkers();
  } catch (std::exception &e) {
    delete L->E;
    delete L;
    return nullptr;
  }
  return L;
}

int vb2_host_summary(void *h, double *avg_depth, double *sd_depth, long *num_bases, int *eff_sites,
                     uint32_t *num_marker, int64_t *n_info, int64_t *n_reads, int *sanity_ok) {
  Loaded *L = static_cast<Loaded *>(h);
  if (!L) return 1;
  const SimplePileupViewer &v = L->E->viewer;
  *avg_depth = v.avgDepth; *sd_depth = v.sdDepth; *num_bases = v.numBases; *eff_sites = v.effectiveNumSite;
  *num_marker = L->E->NumMarker; *n_info = (int64_t)v.NumInfo(); *n_reads = (int64_t)v.bases.size();
  *sanity_ok = L->sanity_ok;
  return 0;
}

int vb2_host_copy(void *h, int32_t *base_info_index, char *alt_base, double *known_af, int64_t *info_offset,
                  char *bases, char *quals, double *ud, double *means) {
  Loaded *L = static_cast<Loaded *>(h);
  if (!L) return 1;
  ContaminationEstimator &E = *L->E;
  for (uint32_t i = 0; i < E.NumMarker; ++i) {
    base_info_index[i] = E.resolvedMarkers[i].baseInfoIndex;
    alt_base[i] = E.resolvedMarkers[i].altBase;
    if (known_af) known_af[i] = E.resolvedMarkers[i].knownAFValue;
    for (int k = 0; k < E.numPC; ++k) ud[(size_t)i * E.numPC + k] = E.UD[i][k];
    means[i] = E.means[i];
  }
  memcpy(info_offset, E.viewer.infoOffset.data(), E.viewer.infoOffset.size() * sizeof(int64_t));
  if (!E.viewer.bases.empty()) {
    memcpy(bases, E.viewer.bases.data(), E.viewer.bases.size());
    memcpy(quals, E.viewer.quals.data(), E.viewer.quals.size());
  }
  return 0;
}

void vb2_host_free(void *h) {
  Loaded *L = static_cast<Loaded *>(h);
  if (!L) return;
  delete L->E;
  delete L;
}

}  // extern "C"
