// tools/microbench_fp64.cu -- what does the FP64 pipe of a B200 SM sub-partition sustain for the operand patterns of the
// read loop?  Per iteration every warp runs the 40 fp64 instructions of eat4 (llk_engine.cu) on values kept in registers:
//   KIND 0  x = fma(x, b, a), a/b kernel arguments (one register operand)           -- the round-1 microbenchmark
//   KIND 1  x = fma(y, z, x), three register operands
//   KIND 2  eat4, coefficients in registers (llk_stream_kernel)
//   KIND 3  eat4, C1/C2 as kernel arguments -> uniform operands, C0 in registers (llk_flow_kernel)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Coef { double C0[6], C1[6], C2[6]; };

template <int KIND>
__global__ void __launch_bounds__(128, 8) k_fp64(const __grid_constant__ Coef K, const double *in, double *out, int iters) {
  double acc[6], e[4];
  double c0[6], c1[6], c2[6];
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    acc[p] = 1.0 + 1e-9 * threadIdx.x;
    c0[p] = in[p]; c1[p] = in[6 + p]; c2[p] = in[12 + p];   // (register copies the compiler cannot fold)
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) e[j] = in[20 + j] * (1 + (threadIdx.x & 3));
  double x[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) x[j] = in[j] + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 12; ++j) x[j] = fma(x[j], K.C1[0], K.C0[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = fma(x[j], K.C1[0], K.C0[0]);
    } else if (KIND == 1) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 12; ++j) x[j] = fma(x[(j + 1) % 12], c1[j % 6], x[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = fma(x[(j + 5) % 12], c2[j], x[j]);
    } else {
      const double s01 = e[0] + e[1], t01 = e[0] * e[1], s23 = e[2] + e[3], t23 = e[2] * e[3];
      double g[6], h[6];
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        g[p] = fma(KIND == 2 ? c1[p] : K.C1[p], s01, c0[p]);
        h[p] = fma(KIND == 2 ? c1[p] : K.C1[p], s23, c0[p]);
      }
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        g[p] = fma(KIND == 2 ? c2[p] : K.C2[p], t01, g[p]);
        h[p] = fma(KIND == 2 ? c2[p] : K.C2[p], t23, h[p]);
      }
#pragma unroll
      for (int p = 0; p < 6; ++p) g[p] *= h[p];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[p] *= g[p];
      // rotate the errors so that nothing is loop-invariant
      const double t = e[0]; e[0] = e[1]; e[1] = e[2]; e[2] = e[3]; e[3] = t;
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int p = 0; p < 6; ++p) s += acc[p];
#pragma unroll
  for (int j = 0; j < 12; ++j) s += x[j];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / iters;
}

template <int KIND>
void run(const char *what, int ctas_per_sm) {
  double *d, *in;
  CK(cudaMalloc(&d, 64));
  CK(cudaMalloc(&in, 64 * sizeof(double)));
  double h_in[64];
  for (int i = 0; i < 64; ++i) h_in[i] = 0.5 + 1e-3 * i;
  CK(cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  Coef K;
  for (int p = 0; p < 6; ++p) { K.C0[p] = 0.25 + 0.01 * p; K.C1[p] = 0.5 + 0.01 * p; K.C2[p] = 1.0 - 0.01 * p; }
  k_fp64<KIND><<<148 * ctas_per_sm, 128>>>(K, in, d, 2000);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_fp64<KIND><<<148 * ctas_per_sm, 128>>>(K, in, d, 20000);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double h[2];
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  // whole-kernel view: fp64 warp-instructions per SM sub-partition / elapsed cycles (at the nominal clock)
  const double cyc = ms * 1e-3 * khz * 1e3, per_smsp = 20000.0 * 40.0 * ctas_per_sm;
  printf("%-48s %d warps/SMSP: block 0 %.1f cycles/iter; kernel %.3f ms = %.3f cycles per fp64 warp-instruction per SMSP\n", what,
         ctas_per_sm, h[1], ms, cyc / per_smsp);
  cudaFree(d); cudaFree(in);
}

int main() {
  for (int w : {1, 2, 4, 8}) {
    run<0>("fma(x, const, const)", w);
    run<1>("fma(y, z, x) three register operands", w);
    run<2>("eat4, coefficients in registers", w);
    run<3>("eat4, C1/C2 uniform operands, C0 in registers", w);
  }
  return 0;
}
