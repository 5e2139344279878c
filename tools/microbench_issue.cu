// tools/microbench_issue.cu -- does a non-FP64 instruction issue "for free" next to FP64 work on a B200 SM sub-partition?
// Each warp runs, per iteration, F independent DFMAs and G integer/LDS instructions.  If FP64 only blocks its pipe (2
// cycles per warp-instruction) the time per iteration is max(2F, F+G) per warp slot; if it also blocks the issue port it
// is 2F + G.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_issue tools/microbench_issue.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int F, int G, int KIND>
__global__ void __launch_bounds__(512, 1) k_mix(double *out, int iters, double a, double b, unsigned m) {
  __shared__ double tab[256];
  if (threadIdx.x < 256) tab[threadIdx.x] = a * threadIdx.x;
  __syncthreads();
  double x[F];
  unsigned y[G > 0 ? G : 1];
#pragma unroll
  for (int j = 0; j < F; ++j) x[j] = a + threadIdx.x + j;
#pragma unroll
  for (int j = 0; j < (G > 0 ? G : 1); ++j) y[j] = threadIdx.x * 7u + j;
  double z = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < (F > G ? F : G); ++j) {
      if (j < F) x[j] = fma(x[j], b, a);
      if (j < G) {
        if (KIND == 0) y[j] = (y[j] ^ m) + (y[j] >> 3);          // 2 integer instructions (LOP3/SHF + IADD)
        else { z += tab[(y[j] & 255u)]; y[j] += m; }                // LDS.64 + IADD (+ DADD: counted as FP64!)
      }
    }
  }
  long long t1 = clock64();
  double s = z;
#pragma unroll
  for (int j = 0; j < F; ++j) s += x[j];
  unsigned u = 0;
#pragma unroll
  for (int j = 0; j < (G > 0 ? G : 1); ++j) u ^= y[j];
  if (s == 123.456 || u == 0x12345u) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / iters;
}

template <int F, int G, int KIND>
void run(const char *what, int threads) {
  double *d;
  CK(cudaMalloc(&d, 64));
  k_mix<F, G, KIND><<<148, threads>>>(d, 2000, 1.0, 0.999, 5u);
  CK(cudaDeviceSynchronize());
  k_mix<F, G, KIND><<<148, threads>>>(d, 20000, 1.0, 0.999, 5u);
  CK(cudaDeviceSynchronize());
  double h[2];
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  const int warps_per_smsp = threads / 128;
  printf("%-44s threads %4d: %.1f cycles/iter per warp = %.2f cycles per SMSP per (F=%d fp64 + G=%d x2 int) group\n", what, threads, h[1],
         h[1] / warps_per_smsp, F, G);
  cudaFree(d);
}

int main() {
  for (int threads : {128, 256, 512}) {
    if (threads == 128) { run<12, 0, 0>("12 DFMA", 128); run<12, 6, 0>("12 DFMA + 6x2 int", 128); run<12, 12, 0>("12 DFMA + 12x2 int", 128); run<0 + 1, 12, 0>("1 DFMA + 12x2 int", 128); }
    if (threads == 256) { run<12, 0, 0>("12 DFMA", 256); run<12, 6, 0>("12 DFMA + 6x2 int", 256); run<12, 12, 0>("12 DFMA + 12x2 int", 256); run<1, 12, 0>("1 DFMA + 12x2 int", 256); }
    if (threads == 512) { run<12, 0, 0>("12 DFMA", 512); run<12, 6, 0>("12 DFMA + 6x2 int", 512); run<12, 12, 0>("12 DFMA + 12x2 int", 512); run<1, 12, 0>("1 DFMA + 12x2 int", 512); }
  }
  return 0;
}
