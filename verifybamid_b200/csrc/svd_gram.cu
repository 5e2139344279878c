// svd_gram.cu -- the panel-construction step of `--RefVCF` on a B200 (include/vb2_svd.h, library libvb2svd.so).
// Replaces the centring of ProcessRefVCF and SVDcalculator::ComputeSvdGram (reference SVDcalculator.cpp:402-409, :258-339):
//   center_kernel   mu[m] = mean_j g[m][j],  A[m][j] = g[m][j] - mu[m]           (cpp:402-409; one warp per marker)
//   gram_kernel     G = A^T A, lower tiles only, split over the markers            (cpp:305-306: rankUpdate(A^T))
//   gram_sum_kernel adds the splits in a fixed order and mirrors the triangle       (deterministic: no atomics)
//   cuSOLVER        Ssyevd of the N x N Gram matrix                                 (cpp:309: SelfAdjointEigenSolver; library)
//   ud_kernel       UD = A * PC for the top n_pc eigenvectors                       (cpp:337-338; one warp per marker)
// Single precision throughout, like the reference's Eigen::MatrixXf.  A is M x N with M (markers, 10^5..10^6) >> N
// (samples, ~2,500): the Gram product is the dense contraction (2 M N^2 / 2 flops), everything else streams A once.
// The product runs on the fp32 FMA pipe, not on tensor cores: the centred genotypes are not exact in bf16/tf32 and the
// eigenvectors of G are compared with the reference's fp32 result (a 3xTF32 tcgen05 version is the obvious next step;
// at N = 2,504, M = 100k this kernel takes tens of milliseconds, the eigensolver hundreds).
// There is NO CPU fallback in this file.
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "vb2_svd.h"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define SVD_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      cleanup();                                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? 4 : 3, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    }                                                                                                    \
  } while (0)

// ---- centring: one warp per marker -------------------------------------------------------------------------------
// The genotypes are small integers, so the row sum is exact in fp32 whatever the order of the additions and the mean
// is one rounding -- the same bits as Eigen's rowwise().mean().
template <typename In>
__global__ void __launch_bounds__(256) center_kernel(const In *__restrict__ g, uint32_t M, uint32_t N, uint32_t ld,
                                                     float *__restrict__ A, float *__restrict__ mu, bool subtract) {
  const uint32_t m = blockIdx.x * 8u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (m >= M) return;
  const In *row = g + (size_t)m * N;
  float mean = 0.f;
  if (subtract) {
    float s = 0.f;
    for (uint32_t j = lane; j < N; j += 32) s += (float)row[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    mean = s / (float)N;
    if (lane == 0 && mu) mu[m] = mean;
  }
  float *out = A + (size_t)m * ld;
  for (uint32_t j = lane; j < ld; j += 32) out[j] = j < N ? (float)row[j] - mean : 0.f;
}

// ---- G = A^T A -----------------------------------------------------------------------------------------------------
// CTA (tile, split): the 128 x 128 tile (ti, tj), tj <= ti, of G over the markers of its split.  256 threads, each an
// 8 x 8 block of the tile; a step stages 16 markers x 128 samples of both operand panels in shared memory (float4
// loads along the samples: A is row-major with a leading dimension that is a multiple of 4).
constexpr int kTile = 128, kStep = 16;
__global__ void __launch_bounds__(256, 2) gram_kernel(const float *__restrict__ A, uint32_t M, uint32_t ld, uint32_t n_tile_rows,
                                                       uint32_t markers_per_split, float *__restrict__ P) {
  __shared__ __align__(16) float As[2][kStep][kTile], Bs[2][kStep][kTile];
  // tile index -> (ti, tj) of the lower triangle
  uint32_t t = blockIdx.x, ti = 0;
  while (t > ti) { t -= ti + 1; ++ti; }
  const uint32_t tj = t;
  const uint32_t i0 = ti * kTile, j0 = tj * kTile;
  const uint32_t m_lo = blockIdx.y * markers_per_split, m_hi = min(M, m_lo + markers_per_split);
  const uint32_t tid = threadIdx.x, ty = tid >> 4, tx = tid & 15u;
  const uint32_t lrow = tid >> 5, lcol = (tid & 31u) * 4u;  // this thread's float4 of a staged panel: rows lrow, lrow + 8

  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  auto stage = [&](uint32_t m0, int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t m = m0 + lrow + 8u * h;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (m < m_hi) {
        const float *row = A + (size_t)m * ld;
        if (i0 + lcol < ld) a = *reinterpret_cast<const float4 *>(row + i0 + lcol);
        if (j0 + lcol < ld) b = *reinterpret_cast<const float4 *>(row + j0 + lcol);
      }
      *reinterpret_cast<float4 *>(&As[buf][lrow + 8 * h][lcol]) = a;
      *reinterpret_cast<float4 *>(&Bs[buf][lrow + 8 * h][lcol]) = b;
    }
  };
  int buf = 0;
  if (m_lo < m_hi) stage(m_lo, 0);
  __syncthreads();
  for (uint32_t m0 = m_lo; m0 < m_hi; m0 += kStep) {
    if (m0 + kStep < m_hi) stage(m0 + kStep, buf ^ 1);  // the next panel while this one is multiplied
#pragma unroll
    for (int k = 0; k < kStep; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      *reinterpret_cast<float4 *>(&b[0]) = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8]);
      *reinterpret_cast<float4 *>(&b[4]) = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8 + 4]);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
    __syncthreads();
    buf ^= 1;
  }
  // partial tile of this split: P[split][i][j], leading dimension n_tile_rows * 128
  const uint32_t ldp = n_tile_rows * kTile;
  float *out = P + (size_t)blockIdx.y * ldp * ldp;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float *dst = out + (size_t)(i0 + ty * 8 + r) * ldp + j0 + tx * 8;
    *reinterpret_cast<float4 *>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
  }
}

// G[i][j] = sum over the splits (in split order) for j <= i, mirrored into G[j][i]
__global__ void __launch_bounds__(256) gram_sum_kernel(const float *__restrict__ P, uint32_t n_split, uint32_t ldp, uint32_t N,
                                                       uint32_t ldg, float *__restrict__ G) {
  const uint32_t j = blockIdx.x * 256u + threadIdx.x, i = blockIdx.y;
  if (i >= N || j > i) return;
  float s = 0.f;
  for (uint32_t k = 0; k < n_split; ++k) s += P[(size_t)k * ldp * ldp + (size_t)i * ldp + j];
  G[(size_t)i * ldg + j] = s;
  G[(size_t)j * ldg + i] = s;
}

// ---- UD = A * PC: one warp per marker; eigenvector c of the top ones is E[(N - 1 - c) * lde + 0 .. N) ---------------
template <int KC>
__global__ void __launch_bounds__(256) ud_kernel(const float *__restrict__ A, uint32_t M, uint32_t N, uint32_t ld,
                                                 const float *__restrict__ E, uint32_t lde, uint32_t c0, uint32_t n_pc,
                                                 float *__restrict__ UD) {
  const uint32_t m = blockIdx.x * 8u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (m >= M) return;
  const float *row = A + (size_t)m * ld;
  float acc[KC];
#pragma unroll
  for (int c = 0; c < KC; ++c) acc[c] = 0.f;
  for (uint32_t j = lane; j < N; j += 32) {
    const float a = row[j];
#pragma unroll
    for (int c = 0; c < KC; ++c)
      if (c0 + c < n_pc) acc[c] = fmaf(a, E[(size_t)(N - 1u - (c0 + c)) * lde + j], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0 && c0 + c < n_pc) UD[(size_t)m * n_pc + c0 + c] = v;
  }
}

}  // namespace

extern "C" const char *vb2_svd_last_error(void) { return g_err.c_str(); }

extern "C" int vb2_svd_gram(const vb2_svd_desc *d) {
  if (!d || d->struct_size != sizeof(vb2_svd_desc)) return fail(1, "vb2_svd_gram: null descriptor or wrong struct_size");
  const uint32_t M = d->n_marker, N = d->n_sample, K = d->n_pc;
  if (M == 0 || N == 0) return fail(1, "vb2_svd_gram: empty matrix");
  if (K < 1 || K > std::min(std::min(M, N), (uint32_t)VB2_SVD_MAX_PC))
    return fail(1, "ComputeSvdGram: numPCs must be in [1, min(M, N, VB2_SVD_MAX_PC)]");
  if ((!d->genotype && !d->centered) || !d->ud || !d->pc || !d->singular) return fail(1, "vb2_svd_gram: null pointer");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || d->device < 0 || d->device >= n_dev) {
    cudaGetLastError();
    return fail(2, "vb2_svd_gram: no usable CUDA device (there is no CPU fallback)");
  }
  float *dA = nullptr, *dMu = nullptr, *dP = nullptr, *dG = nullptr, *dW = nullptr, *dWork = nullptr, *dUD = nullptr;
  void *dIn = nullptr;
  int *dInfo = nullptr;
  cusolverDnHandle_t solver = nullptr;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  auto cleanup = [&]() {
    for (void *p : {(void *)dA, (void *)dMu, (void *)dP, (void *)dG, (void *)dW, (void *)dWork, (void *)dUD, dIn, (void *)dInfo})
      if (p) cudaFree(p);
    if (solver) cusolverDnDestroy(solver);
    for (auto &e : ev)
      if (e) cudaEventDestroy(e);
  };
  SVD_CUDA(cudaSetDevice(d->device));
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, d->device);
  const uint32_t ld = (N + 3u) & ~3u;                        // leading dimension of A and G: float4 loads
  const uint32_t n_tile_rows = (N + kTile - 1) / kTile, ldp = n_tile_rows * kTile;
  const uint32_t n_tiles = n_tile_rows * (n_tile_rows + 1) / 2;
  // splits over the markers: enough CTAs for ~two per SM, each split at least 512 markers deep
  uint32_t n_split = std::max(1u, std::min((2u * (uint32_t)sm_count + n_tiles - 1) / n_tiles, (M + 511u) / 512u));
  n_split = std::min(n_split, 64u);
  const uint32_t per_split = ((M + n_split - 1) / n_split + kStep - 1) / kStep * kStep;
  n_split = (M + per_split - 1) / per_split;

  for (auto &e : ev) SVD_CUDA(cudaEventCreate(&e));
  SVD_CUDA(cudaMalloc(&dA, (size_t)M * ld * sizeof(float)));
  const size_t in_bytes = (size_t)M * N * (d->genotype ? sizeof(int8_t) : sizeof(float));
  SVD_CUDA(cudaMalloc(&dIn, in_bytes));
  SVD_CUDA(cudaMalloc(&dMu, (size_t)M * sizeof(float)));
  SVD_CUDA(cudaMalloc(&dP, (size_t)n_split * ldp * ldp * sizeof(float)));
  SVD_CUDA(cudaMalloc(&dG, (size_t)ld * ld * sizeof(float)));
  SVD_CUDA(cudaMalloc(&dW, (size_t)N * sizeof(float)));
  SVD_CUDA(cudaMalloc(&dUD, (size_t)M * K * sizeof(float)));
  SVD_CUDA(cudaMalloc(&dInfo, sizeof(int)));
  SVD_CUDA(cudaMemcpy(dIn, d->genotype ? (const void *)d->genotype : (const void *)d->centered, in_bytes, cudaMemcpyHostToDevice));
  SVD_CUDA(cudaMemset(dG, 0, (size_t)ld * ld * sizeof(float)));

  SVD_CUDA(cudaEventRecord(ev[0]));
  if (d->genotype)
    center_kernel<int8_t><<<(M + 7) / 8, 256>>>(static_cast<const int8_t *>(dIn), M, N, ld, dA, dMu, true);
  else
    center_kernel<float><<<(M + 7) / 8, 256>>>(static_cast<const float *>(dIn), M, N, ld, dA, dMu, false);
  SVD_CUDA(cudaGetLastError());
  SVD_CUDA(cudaEventRecord(ev[1]));
  gram_kernel<<<dim3(n_tiles, n_split), 256>>>(dA, M, ld, n_tile_rows, per_split, dP);
  SVD_CUDA(cudaGetLastError());
  gram_sum_kernel<<<dim3((N + 255) / 256, N), 256>>>(dP, n_split, ldp, N, ld, dG);
  SVD_CUDA(cudaGetLastError());
  SVD_CUDA(cudaEventRecord(ev[2]));

  // eigendecomposition of the symmetric N x N matrix (both triangles are filled; eigenvalues ascending, eigenvector v
  // in dG[v * ld + 0 .. N) afterwards) -- library code, like Eigen's SelfAdjointEigenSolver in the reference
  if (cusolverDnCreate(&solver) != CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    return fail(3, "cusolverDnCreate failed");
  }
  int lwork = 0;
  if (cusolverDnSsyevd_bufferSize(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)N, dG, (int)ld, dW, &lwork) !=
      CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    return fail(3, "cusolverDnSsyevd_bufferSize failed");
  }
  SVD_CUDA(cudaMalloc(&dWork, (size_t)std::max(lwork, 1) * sizeof(float)));
  if (cusolverDnSsyevd(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)N, dG, (int)ld, dW, dWork, lwork, dInfo) !=
      CUSOLVER_STATUS_SUCCESS) {
    cleanup();
    return fail(3, "cusolverDnSsyevd failed");
  }
  int info = 0;
  SVD_CUDA(cudaMemcpy(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost));
  if (info != 0) {
    cleanup();
    return fail(3, "ComputeSvdGram: Gram matrix eigendecomposition failed to converge");  // (cpp:310-312)
  }
  SVD_CUDA(cudaEventRecord(ev[3]));
  for (uint32_t c0 = 0; c0 < K; c0 += 8)
    ud_kernel<8><<<(M + 7) / 8, 256>>>(dA, M, N, ld, dG, ld, c0, K, dUD);
  SVD_CUDA(cudaGetLastError());
  SVD_CUDA(cudaEventRecord(ev[4]));
  SVD_CUDA(cudaEventSynchronize(ev[4]));

  // results: singular values = sqrt(max(0, eigenvalue)) descending (cpp:315-322); PC = the top eigenvectors (cpp:337)
  std::vector<float> w(N), vec((size_t)K * N);
  SVD_CUDA(cudaMemcpy(w.data(), dW, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < N; ++i) d->singular[i] = std::sqrt(std::max(0.0f, w[N - 1 - i]));
  for (uint32_t c = 0; c < K; ++c)
    SVD_CUDA(cudaMemcpy(vec.data() + (size_t)c * N, dG + (size_t)(N - 1 - c) * ld, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost));
  for (uint32_t j = 0; j < N; ++j)
    for (uint32_t c = 0; c < K; ++c) d->pc[(size_t)j * K + c] = vec[(size_t)c * N + j];
  SVD_CUDA(cudaMemcpy(d->ud, dUD, (size_t)M * K * sizeof(float), cudaMemcpyDeviceToHost));
  if (d->mu && d->genotype) SVD_CUDA(cudaMemcpy(d->mu, dMu, (size_t)M * sizeof(float), cudaMemcpyDeviceToHost));
  if (d->timing) {
    cudaEventElapsedTime(&d->timing->center_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&d->timing->gram_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&d->timing->eigen_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&d->timing->ud_ms, ev[3], ev[4]);
    cudaEventElapsedTime(&d->timing->total_ms, ev[0], ev[4]);
  }
  cleanup();
  return 0;
}
