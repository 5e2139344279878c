// host_shim.cpp -- C entry points over the HOST side of the engine (file readers, text-pileup parser,
// marker resolution, depth sanity filter, Nelder-Mead) so that it can be exercised without a GPU
// (tests/test_host.py) and reused from Python (verifybamid_b200/host.py).  Built into libvb2host.so.
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>

#include <thread>
#include <vector>

#include "amoeba.h"
#include "cohort.h"
#include "estimator.h"
#include "svd_panel.h"

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

// Lock-step logic of cohort mode without a GPU: n samples, sample i minimises (over dim variables, from
// start[i*dim..]) the callback fn(user, x, dim) shifted by i -- f_i(x) = fn(x - 0.1*i) -- with AmoebaMinimizer on
// its own thread, every function value served by one CohortCoordinator whose launcher calls fn on the host.
// out_point[n*dim], out_fmin[n], out_cycles[n]; returns the number of coordinator launches (< 0 on error).
long vb2_host_cohort_selftest(double (*fn)(void *, const double *, int), void *user, int n, int dim, const double *start,
                              double ftol, double *out_point, double *out_fmin, long *out_cycles) {
  CohortCoordinator::Launcher launcher = [&](vb2_llk_ctx *const *ctxs, int m, const double *a, const double *,
                                             const double *al, double *out) -> int {
    for (int j = 0; j < m; ++j) {
      const int i = (int)(intptr_t)ctxs[j] - 1;  // the "context" carries the sample number
      std::vector<double> x(a + (size_t)j * dim, a + (size_t)(j + 1) * dim);
      for (double &v : x) v -= 0.1 * i;
      out[j] = fn(user, x.data(), dim) + 0.0 * al[j];
    }
    return VB2_OK;
  };
  CohortCoordinator C(n, dim, launcher);
  struct Via : VectorFunc {
    CohortCoordinator *c;
    int i;
    double Evaluate(const std::vector<double> &v) override {
      return c->Evaluate(i, (vb2_llk_ctx *)(intptr_t)(i + 1), v.data(), v.data(), 0.0);
    }
  };
  std::thread coord([&]() { C.Run(); });
  std::vector<std::thread> workers;
  for (int i = 0; i < n; ++i)
    workers.emplace_back([&, i]() {
      Via f;
      f.c = &C;
      f.i = i;
      AmoebaMinimizer m;
      m.func = &f;
      m.Reset(dim);
      m.point.assign(start + (size_t)i * dim, start + (size_t)(i + 1) * dim);
      out_fmin[i] = m.Minimize(ftol);
      for (int d = 0; d < dim; ++d) out_point[(size_t)i * dim + d] = m.point[d];
      out_cycles[i] = m.cycleCount;
      C.Finish(i);
    });
  for (auto &w : workers) w.join();
  coord.join();
  return C.error.empty() ? C.launches : -1;
}

void vb2_host_free(void *h) {
  Loaded *L = static_cast<Loaded *>(h);
  if (!L) return;
  delete L->E;
  delete L;
}


// SVDcalculator::ReadVcf on the host (svd_panel.cpp): the genotype matrix [markers][samples] (-1..2) and the marker keys.
// Returns a handle (nullptr on failure, message in err[0..err_len)); sizes through n_marker / n_sample.
struct VcfLoaded {
  SVDcalculator calc;
  std::vector<int8_t> genotype;
};
void *vb2_host_read_vcf(const char *path, const char *include_chr_csv, int *n_marker, int *n_sample, char *err, int err_len) {
  VcfLoaded *L = new VcfLoaded();
  try {
    std::unordered_set<std::string> chr;
    std::string csv = include_chr_csv ? include_chr_csv : "", tok;
    size_t b = 0;
    while (b <= csv.size()) {
      const size_t e = csv.find(',', b);
      tok = csv.substr(b, e == std::string::npos ? std::string::npos : e - b);
      if (!tok.empty()) chr.insert(tok);
      if (e == std::string::npos) break;
      b = e + 1;
    }
    L->calc.ReadVcf(path, L->genotype, *n_sample, *n_marker, chr);
    L->calc.numMarker = *n_marker;
    L->calc.numIndividual = *n_sample;
    return L;
  } catch (const std::exception &ex) {
    if (err && err_len > 0) {
      strncpy(err, ex.what(), (size_t)err_len - 1);
      err[err_len - 1] = 0;
    }
    delete L;
    return nullptr;
  }
}
// genotype [n_marker * n_sample] int8, pos [n_marker] int32, ref / alt [n_marker] char
void vb2_host_vcf_copy(void *h, int8_t *genotype, int32_t *pos, char *ref, char *alt) {
  VcfLoaded *L = static_cast<VcfLoaded *>(h);
  memcpy(genotype, L->genotype.data(), L->genotype.size());
  for (size_t i = 0; i < L->calc.pos.size(); ++i) { pos[i] = L->calc.pos[i]; ref[i] = L->calc.refAllele[i]; alt[i] = L->calc.altAllele[i]; }
}
const char *vb2_host_vcf_chrom(void *h, int i) { return static_cast<VcfLoaded *>(h)->calc.chrom[(size_t)i].c_str(); }
void vb2_host_vcf_free(void *h) { delete static_cast<VcfLoaded *>(h); }
}  // extern "C"
