/* vb2_svd.h -- C ABI of the B200 panel-construction step (library libvb2svd.so).
 *
 * Replaces, for one reference panel,
 *     the centring + SVDcalculator::ComputeSvdGram        reference SVDcalculator.cpp:402-409, :258-339
 * i.e. what `--RefVCF` spends its time in once the VCF is parsed:  mu = rowwise mean of the genotype matrix,
 * A = genotype - mu,  G = A^T A (N x N),  eigendecomposition of G,  PC = top eigenvectors,  UD = A * PC,
 * singular values = sqrt(max(0, eigenvalues)) in descending order.  All of it in single precision, like the
 * reference's Eigen::MatrixXf.  
 *
 * Kept out of libvb2llk.so on purpose: it links cuSOLVER (the N x N symmetric eigensolver is plain library code),
 * which the likelihood path must not pay for at load time.
 *
 * Matrices are row-major: genotype[m][j] / centered[m][j] = marker m, sample j (ReadVcf's
 * std::vector<std::vector<char>>, cpp:24); ud[m][c]; pc[j][c].  Column c of ud / pc belongs to the c-th largest
 * singular value; the sign of a column is the eigensolver's (both ud and pc carry the same sign, so ud . pc products
 * do not depend on it -- cpp:319-325). */
#ifndef VB2_SVD_H_
#define VB2_SVD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB2_SVD_MAX_PC 64 /* columns of UD one call returns (the reference's default --NumSVDPCs is 10) */

typedef struct vb2_svd_timing { /* device milliseconds (CUDA events) of the four steps; filled when given */
  float center_ms, gram_ms, eigen_ms, ud_ms, total_ms;
} vb2_svd_timing;

typedef struct vb2_svd_desc {
  uint32_t struct_size;   /* sizeof(vb2_svd_desc) */
  int32_t device;
  uint32_t n_marker;      /* M */
  uint32_t n_sample;      /* N */
  uint32_t n_pc;          /* 1 .. min(M, N, VB2_SVD_MAX_PC) */
  uint32_t pad_;
  const int8_t *genotype; /* [M][N], values -1 (missing: kept as -1, cpp:169-171,397-401), 0, 1, 2; or NULL */
  const float *centered;  /* [M][N] already mean-centred: used when genotype is NULL (ComputeSvdGram's own argument) */
  float *mu;              /* out [M]: rowwise mean (genotype input only; may be NULL) */
  float *ud;              /* out [M][n_pc] */
  float *pc;              /* out [N][n_pc] */
  float *singular;        /* out [N], descending */
  vb2_svd_timing *timing; /* out, may be NULL */
} vb2_svd_desc;

/* 0 on success; 1 invalid argument, 2 no usable device, 3 CUDA / cuSOLVER error, 4 out of memory.  No CPU fallback. */
int vb2_svd_gram(const vb2_svd_desc *desc);
/* Message of the last failing call on this thread; never NULL. */
const char *vb2_svd_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
