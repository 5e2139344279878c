/* oracle/llk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, fp64, CPU restatement of the reference hot path, used ONLY as the checker by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under
 * verifybamid_b200/ may include, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   - the ten known-answer LLK values of SURVEY.md section 8(c),
 *   - the six golden .Ancestry files + two .selfSM files of the reference's ctest suite
 *     (reference CMakeLists.txt:86-147, resource/test/expected/*), and
 *   - oracle/_ref/vb2_ref (the reference's own sources compiled here) when it is present.
 *
 * What is restated (all citations relative to the reference tree):
 *   vb2o_compute_mix_llks  ContaminationEstimator.h:194-314  ComputeMixLLKs, table form
 *   cond_lk / phred        ContaminationEstimator.h:164-177, :65-74
 *   classify_base          ContaminationEstimator.h:180-184
 *   initial_gf             ContaminationEstimator.h:186-192
 *   vb2o_amoeba_*          MathGenMin.cpp:17-25, :313-443 (Reset / Minimize / Amoeba)
 *   evaluate               ContaminationEstimator.h:339-442  FullLLKFunc::Evaluate
 *   vb2o_optimize_llk      ContaminationEstimator.h:316-337 (Initialize, CalculateLLK0) and
 *                          ContaminationEstimator.cpp:88-332 (OptimizeLLK and the six Optimize*)
 *
 * Data layout: the reference keeps viewer.baseInfo / viewer.qualInfo as
 * vector<vector<char>> (SimplePileupViewer.h:59-60,94-95) and resolvedMarkers[i] =
 * {baseInfoIndex, altBase, knownAFValue} (ContaminationEstimator.h:470-475).  Here the nested
 * vectors are one CSR pair (info_offset, bases/quals) and the struct is three parallel arrays;
 * the arithmetic, its order and its types are the reference's.
 */
#include <ctype.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n_marker;                /* NumMarker: rows of the .UD file                          */
  int n_pc;                    /* numPC                                                    */
  const double *ud;            /* UD[i][k] at ud[i*ud_stride + k]                          */
  int ud_stride;
  const double *means;         /* means[i]                                                 */
  const int *base_info_index;  /* resolvedMarkers[i].baseInfoIndex, -1 = marker absent     */
  const char *alt_base;        /* resolvedMarkers[i].altBase                               */
  const double *known_af;      /* resolvedMarkers[i].knownAFValue, or NULL (!isAFknown)    */
  const long long *info_offset;/* CSR over viewer.baseInfo: entry b = [off[b], off[b+1])   */
  const char *bases;           /* concatenated viewer.baseInfo                             */
  const char *quals;           /* concatenated viewer.qualInfo                             */
  int sanity_disabled;         /* isSanityCheckDisabled                                    */
  double avg_depth, sd_depth;  /* viewer.avgDepth, viewer.sdDepth                          */
  int num_thread;              /* numThread (OpenMP)                                       */
} vb2o_problem;

/* ContaminationEstimator.h:164-177: COND_LK[is_error][genotype][base_class] */
static const double COND_LK[2][3][3] = {
    {{1.0, 0.0, 0.0}, {0.5, 0.5, 0.0}, {0.0, 1.0, 0.0}},
    {{0.0, 1.0 / 3.0, 2.0 / 3.0}, {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0}, {1.0 / 3.0, 0.0, 2.0 / 3.0}},
};

/* ContaminationEstimator.h:180-184 */
static inline int classify_base(char base, char altBase) {
  if (base == '.' || base == ',') return 0;
  if (toupper(base) == toupper(altBase)) return 1;
  return 2;
}

/* ContaminationEstimator.h:186-192 (min_af / max_af: :94-95) */
static inline void initial_gf(double AF, double *GF) {
  const double min_af = 0.00005, max_af = 0.99995;
  if (AF < min_af) AF = min_af;
  if (AF > max_af) AF = max_af;
  GF[0] = (1 - AF) * (1 - AF);
  GF[1] = 2 * (AF) * (1 - AF);
  GF[2] = AF * AF;
}

long long vb2o_eval_count = 0;

/* ContaminationEstimator.h:194-314 */
double vb2o_compute_mix_llks(const vb2o_problem *p, const double *tPC1, const double *tPC2, double alpha) {
  ++vb2o_eval_count;
  /* :65-74 getPhredTable */
  double phredTable[94];
  for (int i = 0; i < 94; ++i) phredTable[i] = pow(10.0, i / -10.0);

  /* :213-229 per-eval table of log mixed emissions */
  const double oneMinusAlpha = 1.0 - alpha;
  double logLkTable[3][94][3][3];
  for (int bc = 0; bc < 3; ++bc) {
    const double lkErr[3] = {COND_LK[1][0][bc], COND_LK[1][1][bc], COND_LK[1][2][bc]};
    const double lkNoErr[3] = {COND_LK[0][0][bc], COND_LK[0][1][bc], COND_LK[0][2][bc]};
    for (int q = 0; q < 94; ++q) {
      const double pErr = phredTable[q];
      const double pNoErr = 1.0 - pErr;
      for (int g1 = 0; g1 < 3; ++g1)
        for (int g2 = 0; g2 < 3; ++g2) {
          double val = (alpha * lkErr[g1] + oneMinusAlpha * lkErr[g2]) * pErr +
                       (alpha * lkNoErr[g1] + oneMinusAlpha * lkNoErr[g2]) * pNoErr;
          logLkTable[bc][q][g1][g2] = log(val);
        }
    }
  }

  double sumLLK = 0;
  const int k_pc = p->n_pc;
#ifdef _OPENMP
  omp_set_num_threads(p->num_thread > 0 ? p->num_thread : 1);
#pragma omp parallel for reduction(+ : sumLLK)
#endif
  for (int i = 0; i < p->n_marker; ++i) {
    /* :238-249 skip rules */
    const int idx = p->base_info_index[i];
    if (idx < 0) continue;
    const long long beg = p->info_offset[idx], end = p->info_offset[idx + 1];
    const size_t size = (size_t)(end - beg);
    if (size == 0) continue;
    if (!p->sanity_disabled &&
        (size < (p->avg_depth - 3 * p->sd_depth) || size > (p->avg_depth + 3 * p->sd_depth)))
      continue;

    /* :251-267 allele frequencies */
    double AF1, AF2;
    if (p->known_af) {
      AF1 = AF2 = p->known_af[i];
    } else {
      AF1 = 0.;
      for (int k = 0; k < k_pc; ++k) AF1 += p->ud[(size_t)i * p->ud_stride + k] * tPC1[k];
      AF1 += p->means[i];
      AF1 /= 2.0;
      AF2 = 0.;
      for (int k = 0; k < k_pc; ++k) AF2 += p->ud[(size_t)i * p->ud_stride + k] * tPC2[k];
      AF2 += p->means[i];
      AF2 /= 2.0;
    }

    double markerLK = 0;
    double GF[3], GF2[3];
    initial_gf(AF1, GF);
    initial_gf(AF2, GF2);
    const char altBase = p->alt_base[i];

    /* :285-303 per-read accumulation of table entries */
    const int depth = (int)size;
    double baseLKAccum[3][3] = {{0}};
    for (int j = 0; j < depth; ++j) {
      const int bc = classify_base(p->bases[beg + j], altBase);
      int q = (int)(unsigned char)p->quals[beg + j] - 33;
      if (q < 0) q = 0;
      else if (q > 93) q = 93;
      for (int g1 = 0; g1 < 3; ++g1)
        for (int g2 = 0; g2 < 3; ++g2) baseLKAccum[g1][g2] += logLkTable[bc][q][g1][g2];
    }
    /* :307-311 marginalise over genotype pairs */
    for (int g1 = 0; g1 < 3; ++g1)
      for (int g2 = 0; g2 < 3; ++g2) markerLK += exp(baseLKAccum[g1][g2]) * GF[g1] * GF2[g2];
    if (markerLK > 0) sumLLK += log(markerLK);
  }
  return sumLLK;
}

/* Number of markers / reads that survive the skip rules (:238-249): R_used of SURVEY 8(d). */
void vb2o_used_counts(const vb2o_problem *p, long long *markers, long long *reads) {
  long long m = 0, r = 0;
  for (int i = 0; i < p->n_marker; ++i) {
    const int idx = p->base_info_index[i];
    if (idx < 0) continue;
    const size_t size = (size_t)(p->info_offset[idx + 1] - p->info_offset[idx]);
    if (size == 0) continue;
    if (!p->sanity_disabled &&
        (size < (p->avg_depth - 3 * p->sd_depth) || size > (p->avg_depth + 3 * p->sd_depth)))
      continue;
    ++m;
    r += (long long)size;
  }
  *markers = m;
  *reads = r;
}

/* ------------------------------------------------------------------------------------------
 * Nelder-Mead: MathGenMin.cpp:313-443 (AmoebaMinimizer) with GeneralMinimizer::Reset :17-25.
 * Vector arithmetic follows statgen/MathVector.cpp:123-176 element by element.
 * ---------------------------------------------------------------------------------------- */
#define VB2O_MAXDIM 64
#define ZEPS 3.0e-10   /* statgen/MathConstant.h:34 */
#define FPMAX 1.0e+100 /* statgen/MathConstant.h:36 */

typedef double (*vb2o_func)(void *user, const double *v, int dim);

typedef struct {
  int dim;
  double point[VB2O_MAXDIM];
  double simplex[VB2O_MAXDIM + 1][VB2O_MAXDIM];
  double y[VB2O_MAXDIM + 1];
  double psum[VB2O_MAXDIM], ptry[VB2O_MAXDIM];
  double fmin;
  long cycleCount, cycleMax;
  vb2o_func func;
  void *user;
} vb2o_amoeba;

static double am_f(vb2o_amoeba *a, const double *v) { return a->func(a->user, v, a->dim); }

/* MathGenMin.cpp:425-443 */
static double am_step(vb2o_amoeba *a, int ihi, double factor) {
  const int dim = a->dim;
  double fac = (1.0 - factor) / dim;
  for (int i = 0; i < dim; ++i) a->ptry[i] = fac * a->psum[i];                      /* SetMultiple */
  for (int i = 0; i < dim; ++i) a->ptry[i] += (factor - fac) * a->simplex[ihi][i];  /* AddMultiple */
  double ytry = am_f(a, a->ptry);
  if (ytry < a->y[ihi]) {
    a->y[ihi] = ytry;
    for (int i = 0; i < dim; ++i) a->psum[i] -= a->simplex[ihi][i];
    for (int i = 0; i < dim; ++i) a->simplex[ihi][i] = a->ptry[i];
    for (int i = 0; i < dim; ++i) a->psum[i] += a->simplex[ihi][i];
  }
  return ytry;
}

/* MathGenMin.cpp:326-423.  `point` is the start on entry, the best vertex on convergence. */
double vb2o_amoeba_minimize(vb2o_amoeba *a, double ftol) {
  const int dim = a->dim;
  int i, ilo, ihi, inhi, m, nvertex = dim + 1;
  double rtol, ysave, ytry;
  a->fmin = FPMAX; /* Reset(), :17-25; directions = identity * 1.0 */
  if (dim == 0) return a->fmin = am_f(a, a->point);
  for (i = 0; i < dim; i++) {
    memcpy(a->simplex[i], a->point, sizeof(double) * dim);
    for (int j = 0; j < dim; ++j) a->simplex[i][j] += (i == j ? 1.0 : 0.0);
    a->y[i] = am_f(a, a->simplex[i]);
    if (a->y[i] < a->fmin) a->fmin = a->y[i];
  }
  memcpy(a->simplex[nvertex - 1], a->point, sizeof(double) * dim);
  a->y[nvertex - 1] = am_f(a, a->simplex[nvertex - 1]);
  if (a->y[nvertex - 1] < a->fmin) a->fmin = a->y[nvertex - 1];
  a->cycleCount = nvertex;
  memcpy(a->psum, a->simplex[0], sizeof(double) * dim);
  for (m = 1; m < nvertex; m++)
    for (int j = 0; j < dim; ++j) a->psum[j] += a->simplex[m][j];

  while (1) {
    if (a->y[0] > a->y[1]) { ilo = inhi = 1; ihi = 0; } else { ilo = inhi = 0; ihi = 1; }
    for (i = 2; i < nvertex; i++) {
      if (a->y[i] <= a->y[ilo]) ilo = i;
      else if (a->y[i] > a->y[ihi]) { inhi = ihi; ihi = i; }
      else if (a->y[i] > a->y[inhi]) inhi = i;
    }
    rtol = 2 * fabs(a->y[ihi] - a->y[ilo]) / (fabs(a->y[ihi]) + fabs(a->y[ilo]) + ZEPS);
    if (rtol < ftol) {
      memcpy(a->point, a->simplex[ilo], sizeof(double) * dim);
      return a->fmin = a->y[ilo];
    }
    if (a->cycleCount > a->cycleMax) return DBL_MAX;

    a->cycleCount += 2;
    ytry = am_step(a, ihi, -1.0);
    if (ytry <= a->y[ilo]) {
      am_step(a, ihi, 2.0);
    } else if (ytry >= a->y[inhi]) {
      ysave = a->y[ihi];
      ytry = am_step(a, ihi, 0.5);
      if (ytry >= ysave) {
        for (i = 0; i < nvertex; i++)
          if (i != ilo) {
            for (int j = 0; j < dim; ++j) a->simplex[i][j] += a->simplex[ilo][j];
            for (int j = 0; j < dim; ++j) a->simplex[i][j] *= 0.5;
            a->y[i] = am_f(a, a->simplex[i]);
          }
        a->cycleCount += dim;
        memcpy(a->psum, a->simplex[0], sizeof(double) * dim);
        for (m = 1; m < nvertex; m++)
          for (int j = 0; j < dim; ++j) a->psum[j] += a->simplex[m][j];
      }
    } else {
      a->cycleCount--;
    }
  }
}

/* Generic entry so tests can drive the restated minimiser with any callback. */
double vb2o_amoeba_run(vb2o_func func, void *user, int dim, double *point, double ftol, long *cycles) {
  vb2o_amoeba a;
  memset(&a, 0, sizeof(a));
  a.dim = dim;
  a.cycleMax = 50000; /* MathGenMin.cpp:313-314 */
  a.func = func;
  a.user = user;
  memcpy(a.point, point, sizeof(double) * dim);
  double r = vb2o_amoeba_minimize(&a, ftol);
  memcpy(point, a.point, sizeof(double) * dim);
  if (cycles) *cycles = a.cycleCount;
  return r;
}

/* ------------------------------------------------------------------------------------------
 * FullLLKFunc state + Evaluate (ContaminationEstimator.h:76-115, :339-442) and the
 * OptimizeLLK driver (ContaminationEstimator.cpp:88-332).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  /* inputs (main.cpp:285-319) */
  int is_heter, is_pc_fixed, is_alpha_fixed;
  double alpha;                  /* Estimator.alpha: 0.5 or --FixAlpha                  */
  double pc_fixed[VB2O_MAXDIM];  /* Estimator.PC[1]: zeros or --FixPC                   */
  double epsilon;
  /* outputs */
  double global_pc[VB2O_MAXDIM], global_pc2[VB2O_MAXDIM], global_alpha, llk1, llk0;
  long long evals;
} vb2o_model;

typedef struct {
  const vb2o_problem *p;
  vb2o_model *m;
  int is_heter; /* toggled by the two-phase schedule, like Estimator.isHeter */
  double fixPC[VB2O_MAXDIM], fixPC2[VB2O_MAXDIM], fixAlpha;
} fn_state;

static double inv_logit(double x) { double e = exp(x); return e / (1. + e); }  /* h:119-122 */
static double logit(double x) { return log(x / (1. - x)); }                     /* h:124-127 */

/* ContaminationEstimator.h:339-442 */
static double evaluate(void *user, const double *v, int dim) {
  fn_state *s = (fn_state *)user;
  vb2o_model *m = s->m;
  const int k = s->p->n_pc;
  double smLLK = 0;
  double tmpPC[VB2O_MAXDIM], tmpPC2[VB2O_MAXDIM];
  if (!s->is_heter) {
    if (m->is_pc_fixed) {
      double tmpAlpha = inv_logit(v[0]);
      smLLK = 0 - vb2o_compute_mix_llks(s->p, s->fixPC, s->fixPC2, tmpAlpha);
      if (smLLK < m->llk1) { m->llk1 = smLLK; m->global_alpha = tmpAlpha; }
    } else if (m->is_alpha_fixed) {
      for (int i = 0; i < k; ++i) tmpPC[i] = v[i];
      smLLK = 0 - vb2o_compute_mix_llks(s->p, tmpPC, tmpPC, s->fixAlpha);
      if (smLLK < m->llk1) {
        m->llk1 = smLLK;
        memcpy(m->global_pc, tmpPC, sizeof(double) * k);
        memcpy(m->global_pc2, tmpPC, sizeof(double) * k);
      }
    } else {
      for (int i = 0; i < k; ++i) tmpPC[i] = v[i];
      double tmpAlpha = inv_logit(v[k]);
      smLLK = 0 - vb2o_compute_mix_llks(s->p, tmpPC, tmpPC, tmpAlpha);
      if (smLLK < m->llk1) {
        m->llk1 = smLLK;
        memcpy(m->global_pc, tmpPC, sizeof(double) * k);
        memcpy(m->global_pc2, tmpPC, sizeof(double) * k);
        m->global_alpha = tmpAlpha;
      }
    }
  } else {
    if (m->is_pc_fixed) {
      for (int i = 0; i < k; ++i) tmpPC[i] = v[i];
      double tmpAlpha = inv_logit(v[k]);
      smLLK = 0 - vb2o_compute_mix_llks(s->p, tmpPC, s->fixPC2, tmpAlpha);
      if (smLLK < m->llk1) {
        m->llk1 = smLLK;
        memcpy(m->global_pc, tmpPC, sizeof(double) * k);
        m->global_alpha = tmpAlpha;
      }
    } else if (m->is_alpha_fixed) {
      for (int i = 0; i < k; ++i) { tmpPC[i] = 0.; tmpPC2[i] = 0.; }
      for (int i = 0; i < dim; ++i) {
        if (i < k) tmpPC[i] = v[i];
        else if (i < 2 * k) tmpPC2[i - k] = v[i];
      }
      smLLK = 0 - vb2o_compute_mix_llks(s->p, tmpPC, tmpPC2, s->fixAlpha);
      if (smLLK < m->llk1) {
        m->llk1 = smLLK;
        memcpy(m->global_pc, tmpPC, sizeof(double) * k);
        memcpy(m->global_pc2, tmpPC2, sizeof(double) * k);
      }
    } else {
      double tmpAlpha = 0.;
      for (int i = 0; i < k; ++i) { tmpPC[i] = 0.; tmpPC2[i] = 0.; }
      for (int i = 0; i < dim; ++i) {
        if (i < k) tmpPC[i] = v[i];
        else if (i < 2 * k) tmpPC2[i - k] = v[i];
        else if (i == 2 * k) tmpAlpha = inv_logit(v[i]);
      }
      smLLK = 0 - vb2o_compute_mix_llks(s->p, tmpPC, tmpPC2, tmpAlpha);
      if (smLLK < m->llk1) {
        m->llk1 = smLLK;
        memcpy(m->global_pc, tmpPC, sizeof(double) * k);
        memcpy(m->global_pc2, tmpPC2, sizeof(double) * k);
        m->global_alpha = tmpAlpha;
      }
    }
  }
  return smLLK;
}

static void run_amoeba(fn_state *s, int dim, double *point) {
  vb2o_amoeba_run(evaluate, s, dim, point, s->m->epsilon, NULL);
}

/* ContaminationEstimator.cpp:88-190 with the Optimize* bodies of :192-332 inlined. */
int vb2o_optimize_llk(const vb2o_problem *p, vb2o_model *m) {
  const int k = p->n_pc;
  if (2 * k + 1 > VB2O_MAXDIM) return -1;
  fn_state s;
  memset(&s, 0, sizeof(s));
  s.p = p;
  s.m = m;
  s.is_heter = m->is_heter;
  double PC[2][VB2O_MAXDIM];
  double alpha = m->alpha;
  for (int i = 0; i < k; ++i) { PC[0][i] = 0.; PC[1][i] = m->pc_fixed[i]; }
  const long long evals0 = vb2o_eval_count;

  /* Initialize(), ContaminationEstimator.h:316-332 */
  for (int i = 0; i < k; ++i) m->global_pc[i] = s.fixPC[i] = m->global_pc2[i] = s.fixPC2[i] = PC[1][i];
  m->global_alpha = s.fixAlpha = alpha;
  m->llk1 = (0 - vb2o_compute_mix_llks(p, s.fixPC, s.fixPC2, s.fixAlpha));
  for (int i = 0; i < k; ++i) PC[0][i] = PC[1][i] = 0.01;
  alpha = 0.03;

  double pt[VB2O_MAXDIM];
  if (!s.is_heter) {
    if (m->is_pc_fixed) { /* OptimizeHomoFixedPC :315-332 */
      pt[0] = logit(alpha);
      run_amoeba(&s, 1, pt);
      alpha = inv_logit(pt[0]);
    } else if (m->is_alpha_fixed) { /* OptimizeHomoFixedAlpha :291-313 */
      for (int i = 0; i < k; ++i) pt[i] = PC[0][i];
      run_amoeba(&s, k, pt);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
    } else { /* OptimizeHomo :265-289 */
      for (int i = 0; i < k; ++i) pt[i] = PC[0][i];
      pt[k] = logit(alpha);
      run_amoeba(&s, k + 1, pt);
      alpha = inv_logit(pt[k]);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
    }
  } else {
    if (m->is_pc_fixed) { /* OptimizeHeterFixedPC = OptimizeHomo :261-263 */
      for (int i = 0; i < k; ++i) pt[i] = PC[0][i];
      pt[k] = logit(alpha);
      run_amoeba(&s, k + 1, pt);
      alpha = inv_logit(pt[k]);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
    } else if (m->is_alpha_fixed) { /* :117-130 */
      s.is_heter = 0;
      for (int i = 0; i < k; ++i) pt[i] = PC[0][i];
      run_amoeba(&s, k, pt);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
      for (int i = 0; i < k; ++i) PC[1][i] = PC[0][i];
      memcpy(m->global_pc2, m->global_pc, sizeof(double) * k);
      s.is_heter = 1;
      /* OptimizeHeterFixedAlpha :228-259 */
      for (int i = 0; i < 2 * k; ++i) pt[i] = i < k ? PC[0][i] : PC[1][i - k];
      run_amoeba(&s, 2 * k, pt);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
      for (int i = k; i < 2 * k; ++i) PC[1][i - k] = pt[i];
    } else { /* :131-145 */
      s.is_heter = 0;
      for (int i = 0; i < k; ++i) pt[i] = PC[0][i];
      pt[k] = logit(alpha);
      run_amoeba(&s, k + 1, pt);
      alpha = inv_logit(pt[k]);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
      for (int i = 0; i < k; ++i) PC[1][i] = PC[0][i];
      memcpy(m->global_pc2, m->global_pc, sizeof(double) * k);
      s.is_heter = 1;
      /* OptimizeHeter :192-226 */
      for (int i = 0; i < 2 * k; ++i) pt[i] = i < k ? PC[0][i] : PC[1][i - k];
      pt[2 * k] = logit(alpha);
      run_amoeba(&s, 2 * k + 1, pt);
      alpha = inv_logit(pt[2 * k]);
      for (int i = 0; i < k; ++i) PC[0][i] = pt[i];
      for (int i = k; i < 2 * k; ++i) PC[1][i - k] = pt[i];
    }
    /* :146-149 (touches PC1 and PC2 only; the reference indexes [1] unconditionally) */
    if (m->global_alpha >= 0.5) {
      double t = m->global_pc[0]; m->global_pc[0] = m->global_pc2[0]; m->global_pc2[0] = t;
      if (k > 1) { t = m->global_pc[1]; m->global_pc[1] = m->global_pc2[1]; m->global_pc2[1] = t; }
    }
  }
  /* CalculateLLK0, ContaminationEstimator.h:334-337 */
  m->llk0 = (0 - vb2o_compute_mix_llks(p, m->global_pc, m->global_pc, 0));
  m->evals = vb2o_eval_count - evals0;
  return 0;
}
