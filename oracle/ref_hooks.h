/* Force-included (after <omp.h>) when oracle/_ref compiles the reference's
 * ContaminationEstimator.h so the driver can count LLK evaluations without touching
 * the reference source: ComputeMixLLKs calls omp_set_num_threads() exactly once per
 * evaluation (ContaminationEstimator.h:233), so that call is routed through a counter. */
#ifndef VB2_ORACLE_REF_HOOKS_H
#define VB2_ORACLE_REF_HOOKS_H
#include <omp.h>
extern long g_vb2_ref_evals;
static inline void vb2_ref_count_eval(int n) { ++g_vb2_ref_evals; omp_set_num_threads(n); }
#define omp_set_num_threads vb2_ref_count_eval
#endif
