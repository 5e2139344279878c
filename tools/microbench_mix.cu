// tools/microbench_mix.cu -- do other instructions issue "for free" next to FP64 instructions on a B200 SM sub-partition,
// and does it depend on how many REGISTER operands the FP64 instruction reads?  Per iteration a warp runs F = 12
// independent DFMAs and G integer instructions (LOP3 / IADD3 on registers); eight warps per sub-partition.
//   NR = 1  x = fma(x, cb, ca)       one register operand, two from the constant bank   (round-1 microbenchmark)
//   NR = 2  x = fma(x, cb, y)        two register operands                              (llk_flow_kernel's read loop)
//   NR = 3  x = fma(x, z, y)         three register operands                            (llk_stream_kernel's read loop)
// time/iteration = max(pipe, F + G) if only the FP64 pipe is held, pipe + G if the issue/operand path is held as well.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_mix tools/microbench_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int NR, int G>
__global__ void __launch_bounds__(128, 8) k_mix(double *out, const double *in, int iters, double ca, double cb, unsigned m) {
  double x[12], y[6], z[4];
  unsigned u[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    x[j] = in[j] + threadIdx.x; y[j % 6] = in[12 + j] * 1e-3; z[j % 4] = in[24 + j];
    u[j] = threadIdx.x * 7u + j;
  }
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      if (NR == 1) x[j] = fma(x[j], cb, ca);
      if (NR == 2) x[j] = fma(x[j], cb, y[j % 6]);
      if (NR == 3) x[j] = fma(x[j], z[j % 4], y[j % 6]);
#pragma unroll
      for (int g = 0; g < G; ++g) u[j] = (u[j] ^ m) + u[(j + 1 + g) % 12];   // LOP3 + IADD (register operands)
    }
  }
  double s = 0;
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < 12; ++j) { s += x[j]; v ^= u[j]; }
  if (s == 123.456 || v == 0x12345u) out[0] = s;
}

template <int NR, int G>
void run(const double *in, double *d) {
  const int iters = 20000;
  k_mix<NR, G><<<148 * 8, 128>>>(d, in, 200, 1.0, 0.999, 5u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  k_mix<NR, G><<<148 * 8, 128>>>(d, in, iters, 1.0, 0.999, 5u);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * khz * 1e3 / (8.0 * iters);   // cycles per iteration of one warp's group, per SMSP
  printf("DFMA with %d register operand(s): 12 DFMA + %2d int per group -> %.1f cycles per group per SMSP\n", NR, 2 * 12 * G, cyc);
}

int main() {
  double *d, *in;
  CK(cudaMalloc(&d, 64));
  CK(cudaMalloc(&in, 64 * sizeof(double)));
  double h_in[64];
  for (int i = 0; i < 64; ++i) h_in[i] = 0.5 + 1e-3 * i;
  CK(cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
  run<1, 0>(in, d); run<1, 1>(in, d); run<1, 2>(in, d);
  run<2, 0>(in, d); run<2, 1>(in, d); run<2, 2>(in, d);
  run<3, 0>(in, d); run<3, 1>(in, d); run<3, 2>(in, d);
  return 0;
}
