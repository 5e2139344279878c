// tools/microbench.cu -- small measurements that size the LLK kernel design on B200 (not product code):
//   1. fixed cost of a kernel launch as a function of CTA shape, dynamic shared memory and argument size
//   2. FP64 DFMA/DMUL latency (dependent chain) and throughput (independent chains) per SM sub-partition
//   3. cost of the result paths: last-block reduction vs per-CTA stores into host-mapped memory
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Big { double pad[256]; };   // 2 KiB of arguments
struct Small { double pad[4]; };

template <typename P>
__global__ void k_empty(const __grid_constant__ P p, double *out) {
  extern __shared__ char smem[];
  if (threadIdx.x == 0 && p.pad[0] == 123.456) out[blockIdx.x] = smem[0];
}

// touches its arguments the way the LLK kernel does (uniform reads) and syncs once
template <typename P>
__global__ void k_touch(const __grid_constant__ P p, double *out) {
  __shared__ double s[32];
  double v = 0;
  for (int i = 0; i < (int)(sizeof(P) / 8); i += 8) v += p.pad[i];
  if (threadIdx.x < 32) s[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0 && s[3] == 123.456) out[blockIdx.x] = v;
}

__global__ void k_fp64_chain(double *out, int iters, double a, double b) {
  // one dependent chain per thread: latency
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, b, a);
  }
  long long t1 = clock64();
  if (x == 123.456) out[0] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / (iters * 16.0);
}

template <int ILP>
__global__ void k_fp64_tput(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = a + threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 123.456) out[0] = s;
  // cycles per warp-instruction per SM sub-partition = elapsed / (instr per warp * warps per SMSP)
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = (double)(t1 - t0) / (iters * 4.0 * ILP);
}

// result paths
__global__ void k_lastblock(double *partials, unsigned *ticket, double *out, volatile unsigned long long *host_slot,
                            unsigned long long seq) {
  __shared__ int last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = 1.0;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double s = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) s += __ldcg(partials + i);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
    if (threadIdx.x == 0) {
      *ticket = 0;
      out[0] = s;
      if (host_slot) *(ulonglong2 *)host_slot = make_ulonglong2((unsigned long long)__double_as_longlong(s), seq);
    }
  }
}
__global__ void k_hostslots(volatile unsigned long long *host_slots, unsigned long long seq) {
  if (threadIdx.x == 0)
    *(ulonglong2 *)(host_slots + 2 * blockIdx.x) = make_ulonglong2((unsigned long long)__double_as_longlong(1.0), seq);
}

template <typename F>
float time_launches(F launch, int n = 2000) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 50; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < n; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / n;  // us per launch
}

int main() {
  double *d_out;
  CK(cudaMalloc(&d_out, 4096 * 8));
  Big big; memset(&big, 0, sizeof(big));
  Small small; memset(&small, 0, sizeof(small));
  CK(cudaFuncSetAttribute(k_empty<Big>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_empty<Small>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("== back-to-back launch cost (us per kernel, CUDA events over 2000 launches) ==\n");
  struct Cfg { int grid, block, smem; };
  Cfg cfgs[] = {{148, 768, 0}, {148, 768, 72 * 1024}, {148, 768, 150 * 1024}, {148, 256, 0}, {148, 128, 0},
                {296, 384, 36 * 1024}, {444, 256, 24 * 1024}, {592, 128, 0}, {592, 192, 18 * 1024}, {1184, 64, 0},
                {3118, 32, 0}, {780, 128, 8 * 1024}};
  for (auto c : cfgs) {
    float a = time_launches([&] { k_empty<Small><<<c.grid, c.block, c.smem>>>(small, d_out); });
    float b = time_launches([&] { k_empty<Big><<<c.grid, c.block, c.smem>>>(big, d_out); });
    float t = time_launches([&] { k_touch<Big><<<c.grid, c.block, 0>>>(big, d_out); });
    printf("grid %5d block %4d smem %6d : empty(32B args) %6.2f  empty(2KiB args) %6.2f  touch-args+sync(2KiB) %6.2f\n",
           c.grid, c.block, c.smem, a, b, t);
  }

  printf("== FP64 pipe (one CTA on one SM) ==\n");
  double h[2];
  k_fp64_chain<<<1, 32>>>(d_out, 1000, 1.0, 0.999);
  CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
  printf("DFMA dependent-chain latency: %.2f cycles\n", h[1]);
  for (int warps : {1, 2, 4, 6, 8}) {
    // warps per SMSP = warps (block = 4*warps warps)
    k_fp64_tput<1><<<1, 128 * warps>>>(d_out, 2000, 1.0, 0.999); CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost)); double i1 = h[1] / warps;
    k_fp64_tput<2><<<1, 128 * warps>>>(d_out, 2000, 1.0, 0.999); CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost)); double i2 = h[1] / warps;
    k_fp64_tput<6><<<1, 128 * warps>>>(d_out, 2000, 1.0, 0.999); CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost)); double i6 = h[1] / warps;
    k_fp64_tput<12><<<1, 128 * warps>>>(d_out, 2000, 1.0, 0.999); CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost)); double i12 = h[1] / warps;
    printf("%d warps/SMSP: cycles per DFMA warp-instr per SMSP  ILP1 %.2f  ILP2 %.2f  ILP6 %.2f  ILP12 %.2f\n", warps, i1, i2, i6, i12);
  }

  printf("== result paths (grid 148 x 128 threads; us per kernel back-to-back, and host-visible latency) ==\n");
  double *partials; unsigned *ticket;
  CK(cudaMalloc(&partials, 148 * 8)); CK(cudaMalloc(&ticket, 4)); CK(cudaMemset(ticket, 0, 4));
  unsigned long long *h_slots, *d_slots;
  CK(cudaHostAlloc((void **)&h_slots, 16 * 256, cudaHostAllocMapped));
  memset(h_slots, 0, 16 * 256);
  CK(cudaHostGetDevicePointer((void **)&d_slots, h_slots, 0));
  unsigned long long seq = 0;
  printf("last-block, device out only      : %6.2f\n", time_launches([&] { k_lastblock<<<148, 128>>>(partials, ticket, d_out, nullptr, 0); }));
  printf("last-block + one host slot       : %6.2f\n", time_launches([&] { k_lastblock<<<148, 128>>>(partials, ticket, d_out, d_slots, ++seq); }));
  printf("148 host slots (no device reduce): %6.2f\n", time_launches([&] { k_hostslots<<<148, 128>>>(d_slots, ++seq); }));
  // host-visible round trip: launch, spin until the slot(s) carry seq
  auto rt = [&](int mode) {
    const int n = 2000;
    CK(cudaDeviceSynchronize());
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) {
      ++seq;
      if (mode == 0) { k_lastblock<<<148, 128>>>(partials, ticket, d_out, d_slots, seq); while (((volatile unsigned long long *)h_slots)[1] != seq) {} }
      else if (mode == 1) { k_hostslots<<<148, 128>>>(d_slots, seq); for (int c = 0; c < 148; ++c) while (((volatile unsigned long long *)h_slots)[2 * c + 1] != seq) {} }
      else { k_lastblock<<<148, 128>>>(partials, ticket, d_out, nullptr, 0); CK(cudaStreamSynchronize(0)); }
    }
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / n;
  };
  printf("round trip launch->host sees result: last-block+slot spin %.2f us | 148 slots spin %.2f us | streamSynchronize %.2f us\n",
         rt(0), rt(1), rt(2));
  return 0;
}
