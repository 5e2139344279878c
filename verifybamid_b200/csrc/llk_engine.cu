// llk_engine.cu -- the sm_100a contamination-likelihood kernel and the C ABI around it
// (include/vb2_llk.h).  Replaces, for one sample resident in HBM,
//     FullLLKFunc::ComputeMixLLKs            reference ContaminationEstimator.h:194-314
// One evaluation = one launch of llk_kernel:
//   (i)   AF = (UD.PC + mu)/2 per marker (h:251-267), coalesced column-major panel reads,
//         Hardy-Weinberg genotype priors (h:186-192);
//   (ii)  per read, the six alpha-dependent genotype-pair emissions of the 3x3 mixture
//         (h:213-229, in the closed form of SURVEY.md Appendix A: each is LINEAR in the Phred
//         error e, F_p(e) = c0_p + c1_p*e, so one DFMA forms it and one DMUL accumulates it);
//         the read tile of each warp is staged into shared memory by one TMA bulk copy
//         (cp.async.bulk + mbarrier);
//   (iii) log of the marginal per marker (h:307-311), fixed-order warp-shuffle / block / grid
//         reduction in fp64 (h:232-236 is an OpenMP reduction) -> one double.
// Everything that is evaluation-invariant was folded at create time by llk_pack.cpp.
//
// There is NO CPU fallback in this file: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "llk_pack.h"
#include "vb2_llk.h"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxArgJobs = 8;     // evaluations whose parameters travel in the kernel arguments
constexpr int kStageWordsCap = 64; // words per lane per shared-memory stage (256 reads)
constexpr int kNumPairs = 6;       // off-diagonal genotype pairs

// Pair p = (g1 contaminant, g2 intended): 0:(0,1) 1:(0,2) 2:(1,0) 3:(1,2) 4:(2,0) 5:(2,1).
// The alt-allele emission is the ref-allele one with g -> 2-g (COND_LK, h:164-177), and
// (g1,g2) -> (2-g1,2-g2) maps pair p to pair 5-p: an alt read multiplies acc[5-p] by F_p(e).
__host__ __device__ constexpr int pair_g1(int p) { return p < 2 ? 0 : (p < 4 ? 1 : 2); }
__host__ __device__ constexpr int pair_g2(int p) {
  return p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 2 : p == 4 ? 0 : 1;
}

struct JobParams {  // one evaluation (352 bytes)
  double c0[kNumPairs], c1[kNumPairs];
  double pc1[VB2_MAX_PC], pc2[VB2_MAX_PC];  // contaminant / intended PCs
};

struct SampleDev {  // one sample resident in HBM
  const uint32_t *words;
  const uint2 *slice_desc;
  const void *ud;          // [n_pc][m_pad] float or double
  const void *mu;          // [m_pad]
  const double *diag;      // [3][m_pad]
  const double *known_af;  // [m_pad] or nullptr
  const double *phred;     // [128]; entries >= 94 unused
  double *partials;        // [slots][grid_x]
  unsigned int *tickets;   // [slots]
  double log_other_const, min_af, max_af;
  uint32_t n_used, n_slices, m_pad, n_pc;
  uint32_t grid_x, panel_fp64, stage_words, n_buf;
};

struct Mailbox {  // host-mapped, written by the last CTA of a launch
  volatile unsigned long long seq;
  unsigned long long pad_[7];
  volatile double val[VB2_MAX_BATCH];
};

struct LaunchArgs {
  SampleDev sample;           // used when samples == nullptr
  const SampleDev *samples;   // eval_many: job j evaluates samples[j]
  const uint32_t *slots;      // eval_many: partial/ticket slot of job j inside its sample
  const JobParams *jobs_dev;  // parameters in HBM (n_jobs > kMaxArgJobs or eval_many)
  double *d_out;              // [n_jobs] device results (may be nullptr)
  Mailbox *mbox;              // device view of the host mailbox (may be nullptr)
  unsigned int *jobs_done;    // second-level ticket
  unsigned long long seq;
  uint32_t n_jobs;
  uint32_t pad_;
  JobParams jobs[kMaxArgJobs];
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Four reads (one word) of one lane.  ALT = the word holds alt-allele reads.
template <bool ALT>
__device__ __forceinline__ void eat_word(uint32_t w, const double *s_e, const double (&c0)[kNumPairs],
                                         const double (&c1)[kNumPairs], double (&acc)[kNumPairs]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint32_t q = (w >> (8 * b)) & 0xFFu;
    if (q != 0xFFu) {
      const double e = s_e[q];
#pragma unroll
      for (int p = 0; p < kNumPairs; ++p) {
        const double f = fma(c1[p], e, c0[p]);
        acc[ALT ? (kNumPairs - 1 - p) : p] *= f;
      }
    }
  }
}

// ContaminationEstimator.h:186-192 with the reference's comparison order (NaN passes through).
__device__ __forceinline__ void initial_gf(double af, double min_af, double max_af, double (&gf)[3]) {
  if (af < min_af) af = min_af;
  if (af > max_af) af = max_af;
  gf[0] = __dmul_rn(1 - af, 1 - af);
  gf[1] = __dmul_rn(__dmul_rn(2, af), 1 - af);
  gf[2] = __dmul_rn(af, af);
}

template <typename PanelT>
__device__ __forceinline__ void marker_af(const SampleDev &S, const JobParams &J, uint32_t pm, double &af1,
                                          double &af2) {
  // h:251-267: AF = (sum_k UD[i][k]*PC[k] + means[i]) / 2, accumulated in k order in fp64.
  const PanelT *ud = static_cast<const PanelT *>(S.ud);
  const PanelT *mu = static_cast<const PanelT *>(S.mu);
  double a1 = 0., a2 = 0.;
  for (uint32_t k = 0; k < S.n_pc; ++k) {
    const double u = (double)__ldg(ud + (size_t)k * S.m_pad + pm);
    a1 = __dadd_rn(a1, __dmul_rn(u, J.pc1[k]));
    a2 = __dadd_rn(a2, __dmul_rn(u, J.pc2[k]));
  }
  const double m = (double)__ldg(mu + pm);
  af1 = (a1 + m) * 0.5;
  af2 = (a2 + m) * 0.5;
}

__global__ void __launch_bounds__(kThreads)
llk_kernel(const __grid_constant__ LaunchArgs A) {
  extern __shared__ __align__(128) uint32_t s_words[];  // [warp][buf][stage_words][32]
  __shared__ double s_e[128];
  __shared__ JobParams s_job;
  __shared__ double s_red[kWarpsPerCta];
  __shared__ __align__(8) uint64_t s_bar[kWarpsPerCta][2];
  __shared__ SampleDev s_sample;
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.y;

  // ---- per-CTA set-up: which sample, which parameters ----------------------------------------
  if (A.samples) {
    if (threadIdx.x < sizeof(SampleDev) / 8)
      reinterpret_cast<uint64_t *>(&s_sample)[threadIdx.x] =
          reinterpret_cast<const uint64_t *>(A.samples + job)[threadIdx.x];
  } else {
    if (threadIdx.x < sizeof(SampleDev) / 8)
      reinterpret_cast<uint64_t *>(&s_sample)[threadIdx.x] =
          reinterpret_cast<const uint64_t *>(&A.sample)[threadIdx.x];
  }
  if (threadIdx.x < sizeof(JobParams) / 8) {
    const double *src = A.jobs_dev ? reinterpret_cast<const double *>(A.jobs_dev + job)
                                   : reinterpret_cast<const double *>(&A.jobs[job < kMaxArgJobs ? job : 0]);
    reinterpret_cast<double *>(&s_job)[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  const SampleDev &S = s_sample;
  if (blockIdx.x >= S.grid_x) return;  // eval_many: this sample needs fewer CTAs than the grid has
  s_e[threadIdx.x] = __ldg(S.phred + threadIdx.x);  // kThreads == 128 entries
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint32_t slot = A.slots ? A.slots[job] : job;
  const uint32_t slice = blockIdx.x * kWarpsPerCta + warp;
  double v = 0.0;
  if (slice < S.n_slices) {
    const uint2 sd = S.slice_desc[slice];
    const uint32_t wr = sd.y & 0xFFFFu, wa = sd.y >> 16, W = wr + wa;
    const uint32_t stage_words = S.stage_words;
    uint32_t *buf0 = s_words + (size_t)warp * S.n_buf * stage_words * 32;
    const uint32_t *gsrc = S.words + sd.x;

    // ---- (ii-a) kick off the first read tile: one TMA bulk copy per warp -----------------------
    uint32_t n0 = W < stage_words ? W : stage_words;
    if (lane == 0 && n0) {
      mbar_arrive_expect_tx(&s_bar[warp][0], n0 * 128u);
      bulk_g2s(buf0, gsrc, n0 * 128u, &s_bar[warp][0]);
    }

    // ---- (i) allele frequencies and genotype priors -------------------------------------------
    const uint32_t pm = slice * 32 + lane;
    const bool valid = pm < S.n_used;
    double af1, af2;
    if (S.known_af) {
      af1 = af2 = __ldg(S.known_af + pm);  // h:251-252
    } else if (S.panel_fp64) {
      marker_af<double>(S, s_job, pm, af1, af2);
    } else {
      marker_af<float>(S, s_job, pm, af1, af2);
    }
    double gf[3], gf2[3];
    initial_gf(af1, S.min_af, S.max_af, gf);   // contaminating sample
    initial_gf(af2, S.min_af, S.max_af, gf2);  // intended sample
    const double d0 = __ldg(S.diag + pm), d1 = __ldg(S.diag + S.m_pad + pm),
                 d2 = __ldg(S.diag + 2 * (size_t)S.m_pad + pm);

    double c0[kNumPairs], c1[kNumPairs], acc[kNumPairs];
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) {
      c0[p] = s_job.c0[p];
      c1[p] = s_job.c1[p];
      acc[p] = 1.0;
    }

    // ---- (ii-b) stream the read tile(s) ---------------------------------------------------------
    uint32_t parity = 0u;  // bit b = phase of buffer b's mbarrier
    uint32_t b = 0;
    for (uint32_t t0 = 0; t0 < W; t0 += stage_words) {
      const uint32_t n = (W - t0) < stage_words ? (W - t0) : stage_words;
      const uint32_t t1 = t0 + stage_words;  // start of the next stage (only reached when n_buf == 2)
      if (t1 < W) {
        // prefetch the next stage into the other buffer; every lane finished reading it in the
        // previous iteration
        __syncwarp();
        if (lane == 0) {
          const uint32_t nn = (W - t1) < stage_words ? (W - t1) : stage_words;
          mbar_arrive_expect_tx(&s_bar[warp][b ^ 1], nn * 128u);
          bulk_g2s(buf0 + (size_t)(b ^ 1) * stage_words * 32, gsrc + (size_t)t1 * 32, nn * 128u,
                   &s_bar[warp][b ^ 1]);
        }
      }
      mbar_wait(&s_bar[warp][b], (parity >> b) & 1u);
      parity ^= 1u << b;
      const uint32_t *buf = buf0 + (size_t)b * stage_words * 32 + lane;
      const uint32_t n_ref = wr > t0 ? ((wr - t0) < n ? (wr - t0) : n) : 0u;
      uint32_t t = 0;
      for (; t < n_ref; ++t) eat_word<false>(buf[t * 32], s_e, c0, c1, acc);
      for (; t < n; ++t) eat_word<true>(buf[t * 32], s_e, c0, c1, acc);
      b ^= 1u;
    }

    // ---- (iii) marginal over the nine genotype pairs, log ---------------------------------------
    // h:307-311: markerLK = sum_{g1,g2} exp(acc) * GF[g1] * GF2[g2]; here exp(acc) is the running
    // product itself, and the diagonal products are the create-time constants d0..d2.
    double L = d0 * (gf[0] * gf2[0]) + d1 * (gf[1] * gf2[1]) + d2 * (gf[2] * gf2[2]);
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) L += acc[p] * (gf[pair_g1(p)] * gf2[pair_g2(p)]);
    if (valid && L > 0) v = log(L);
  }

  // ---- fixed-order reduction: warp shuffle tree -> CTA -> last CTA sums the grid ---------------
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double cta = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    S.partials[(size_t)slot * S.grid_x + blockIdx.x] = cta;
    __threadfence();
    const unsigned int t = atomicAdd(S.tickets + slot, 1u);
    s_last = (t == S.grid_x - 1);
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    const double *part = S.partials + (size_t)slot * S.grid_x;
    double s = 0.0;
    for (uint32_t i = lane; i < S.grid_x; i += 32) s += __ldcg(part + i);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (lane == 0) {
      S.tickets[slot] = 0u;  // ready for the next launch on this stream
      const double out = s + S.log_other_const;
      if (A.d_out) A.d_out[job] = out;
      if (A.mbox) {
        A.mbox->val[job] = out;
        __threadfence_system();
        const unsigned int done = atomicAdd(A.jobs_done, 1u);
        if (done == A.n_jobs - 1) {
          *A.jobs_done = 0u;
          __threadfence_system();
          A.mbox->seq = A.seq;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct vb2_llk_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool spin = true;
  SampleDev S{};
  SampleDev *d_sample = nullptr;  // device copy of S (for eval_many tables)
  vb2::PackedSample meta;         // sizes only (vectors released after upload)
  std::vector<void *> allocs;
  uint64_t device_bytes = 0;
  uint32_t slots = 0;
  uint32_t smem_bytes = 0;
  Mailbox *h_mbox = nullptr, *d_mbox = nullptr;
  unsigned int *d_jobs_done = nullptr;
  JobParams *h_jobs = nullptr;  // pinned staging [VB2_MAX_BATCH]
  JobParams *d_jobs = nullptr;
  double *d_out = nullptr;      // [VB2_MAX_BATCH]
  // eval_many staging (owned by the leading context)
  SampleDev *h_many = nullptr, *d_many = nullptr;
  uint32_t *h_slots = nullptr, *d_slots = nullptr;
  unsigned long long seq = 0;
  double spin_timeout_ms = 20000.0;
  std::string err;
};

namespace {

int set_err(vb2_llk_ctx *ctx, int code, const std::string &msg) {
  g_last_error = msg;
  if (ctx) ctx->err = msg;
  return code;
}

#define VB2_CUDA(ctx, call)                                                                         \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      return set_err(ctx, e_ == cudaErrorMemoryAllocation ? VB2_ERR_NOMEM : VB2_ERR_CUDA,           \
                     std::string(#call) + ": " + cudaGetErrorString(e_));                           \
    }                                                                                               \
  } while (0)

template <typename T>
int upload(vb2_llk_ctx *ctx, const std::vector<T> &h, const T **d, bool count_bytes = true) {
  *d = nullptr;
  if (h.empty()) return VB2_OK;
  void *p = nullptr;
  VB2_CUDA(ctx, cudaMalloc(&p, h.size() * sizeof(T)));
  ctx->allocs.push_back(p);
  VB2_CUDA(ctx, cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  if (count_bytes) ctx->device_bytes += h.size() * sizeof(T);
  *d = static_cast<const T *>(p);
  return VB2_OK;
}

// Evaluation-dependent constants of the six off-diagonal pairs for ref-class reads:
//   F_p(e) = (alpha*E[g1] + (1-alpha)*E[g2]) * e + (alpha*N[g1] + (1-alpha)*N[g2]) * (1 - e)
//          = c0_p + c1_p * e          (h:223-224 with COND_LK of h:164-177, base class 0)
void fill_job(JobParams *J, uint32_t n_pc, const double *pc1, const double *pc2, double alpha) {
  static const double E[3] = {0.0, 1.0 / 6.0, 1.0 / 3.0};  // COND_LK[1][g][0]
  static const double N[3] = {1.0, 0.5, 0.0};              // COND_LK[0][g][0]
  const double one_minus_alpha = 1.0 - alpha;
  for (int p = 0; p < kNumPairs; ++p) {
    const int g1 = pair_g1(p), g2 = pair_g2(p);
    const double e_mix = alpha * E[g1] + one_minus_alpha * E[g2];
    const double n_mix = alpha * N[g1] + one_minus_alpha * N[g2];
    J->c0[p] = n_mix;
    J->c1[p] = e_mix - n_mix;
  }
  for (uint32_t k = 0; k < VB2_MAX_PC; ++k) {
    J->pc1[k] = k < n_pc ? pc1[k] : 0.0;
    J->pc2[k] = k < n_pc ? pc2[k] : 0.0;
  }
}

int ensure_slots(vb2_llk_ctx *ctx, uint32_t need) {
  if (need <= ctx->slots) return VB2_OK;
  uint32_t n = ctx->slots ? ctx->slots : 8;
  while (n < need) n *= 2;
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double *partials = nullptr;
  unsigned int *tickets = nullptr;
  const size_t gx = ctx->S.grid_x ? ctx->S.grid_x : 1;
  VB2_CUDA(ctx, cudaMalloc(&partials, (size_t)n * gx * sizeof(double)));
  VB2_CUDA(ctx, cudaMalloc(&tickets, (size_t)n * sizeof(unsigned int)));
  VB2_CUDA(ctx, cudaMemsetAsync(tickets, 0, (size_t)n * sizeof(unsigned int), ctx->stream));
  if (ctx->S.partials) cudaFree(ctx->S.partials);
  if (ctx->S.tickets) cudaFree(ctx->S.tickets);
  ctx->S.partials = partials;
  ctx->S.tickets = tickets;
  ctx->slots = n;
  VB2_CUDA(ctx, cudaMemcpyAsync(ctx->d_sample, &ctx->S, sizeof(SampleDev), cudaMemcpyHostToDevice, ctx->stream));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VB2_OK;
}

int wait_mailbox(vb2_llk_ctx *ctx, unsigned long long seq) {
  if (!ctx->spin) {
    VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_mbox->seq != seq) return set_err(ctx, VB2_ERR_CUDA, "kernel finished without publishing its result");
    return VB2_OK;
  }
  // Poll the host-mapped sequence word: cheaper than a stream synchronise for a ~5 us kernel.
  unsigned long spins = 0;
  auto t0 = std::chrono::steady_clock::now();
  while (ctx->h_mbox->seq != seq) {
    if ((++spins & 0x3FFFu) == 0) {
      cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady)
        return set_err(ctx, VB2_ERR_CUDA, std::string("llk_kernel: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && ctx->h_mbox->seq != seq)
        return set_err(ctx, VB2_ERR_CUDA, "kernel finished without publishing its result");
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > ctx->spin_timeout_ms) return set_err(ctx, VB2_ERR_TIMEOUT, "timed out waiting for the device");
    }
  }
  return VB2_OK;
}

// Launch n evaluations of ONE sample.  Results go to d_out (device) and, if to_mailbox, to the
// host mailbox with sequence number *seq_out.
int launch_batch(vb2_llk_ctx *ctx, int n, const double *pc1, const double *pc2, const double *alphas,
                 double *d_out, bool to_mailbox, unsigned long long *seq_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (n <= 0 || n > VB2_MAX_BATCH) return set_err(ctx, VB2_ERR_INVALID, "batch size out of range");
  if (!pc1 || !pc2 || !alphas) return set_err(ctx, VB2_ERR_INVALID, "null parameter array");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_slots(ctx, (uint32_t)n);
  if (rc) return rc;
  LaunchArgs A;
  memset(&A, 0, sizeof(A));
  A.sample = ctx->S;
  A.n_jobs = (uint32_t)n;
  A.d_out = d_out;
  const uint32_t k = ctx->S.n_pc;
  if (n <= kMaxArgJobs) {
    for (int j = 0; j < n; ++j) fill_job(&A.jobs[j], k, pc1 + (size_t)j * k, pc2 + (size_t)j * k, alphas[j]);
  } else {
    // the pinned staging buffer may still be read by the previous batch's copy
    VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < n; ++j) fill_job(&ctx->h_jobs[j], k, pc1 + (size_t)j * k, pc2 + (size_t)j * k, alphas[j]);
    VB2_CUDA(ctx, cudaMemcpyAsync(ctx->d_jobs, ctx->h_jobs, (size_t)n * sizeof(JobParams), cudaMemcpyHostToDevice,
                                  ctx->stream));
    A.jobs_dev = ctx->d_jobs;
  }
  if (to_mailbox) {
    A.mbox = ctx->d_mbox;
    A.jobs_done = ctx->d_jobs_done;
    A.seq = ++ctx->seq;
    if (seq_out) *seq_out = A.seq;
  }
  if (ctx->S.grid_x == 0) return VB2_OK;  // no usable marker: handled by the callers
  dim3 grid(ctx->S.grid_x, (unsigned)n, 1), block(kThreads, 1, 1);
  llk_kernel<<<grid, block, ctx->smem_bytes, ctx->stream>>>(A);
  VB2_CUDA(ctx, cudaGetLastError());
  return VB2_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int vb2_abi_version(void) { return VB2_ABI_VERSION; }

int vb2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *vb2_last_error(const vb2_llk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

void vb2_llk_destroy(vb2_llk_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (void *p : ctx->allocs) cudaFree(p);
  if (ctx->S.partials) cudaFree(ctx->S.partials);
  if (ctx->S.tickets) cudaFree(ctx->S.tickets);
  if (ctx->d_sample) cudaFree(ctx->d_sample);
  if (ctx->d_jobs_done) cudaFree(ctx->d_jobs_done);
  if (ctx->d_jobs) cudaFree(ctx->d_jobs);
  if (ctx->d_out) cudaFree(ctx->d_out);
  if (ctx->d_many) cudaFree(ctx->d_many);
  if (ctx->d_slots) cudaFree(ctx->d_slots);
  if (ctx->h_mbox) cudaFreeHost(ctx->h_mbox);
  if (ctx->h_jobs) cudaFreeHost(ctx->h_jobs);
  if (ctx->h_many) cudaFreeHost(ctx->h_many);
  if (ctx->h_slots) cudaFreeHost(ctx->h_slots);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int create_impl(const vb2_llk_desc *desc, vb2_llk_ctx *ctx) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(ctx, VB2_ERR_NO_DEVICE, "no CUDA device available (this engine has no CPU fallback)");
  }
  if (desc->device < 0 || desc->device >= ndev) return set_err(ctx, VB2_ERR_NO_DEVICE, "desc.device out of range");
  ctx->device = desc->device;
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  VB2_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  if (prop.major < 10)
    return set_err(ctx, VB2_ERR_NO_DEVICE, std::string("device is not Blackwell (sm_100a) : ") + prop.name);
  ctx->sm_count = prop.multiProcessorCount;
  if (desc->stream) {
    ctx->stream = static_cast<cudaStream_t>(desc->stream);
  } else {
    VB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  ctx->spin = !(desc->flags & VB2_FLAG_NO_SPIN);
  if (const char *t = getenv("VB2_LLK_SPIN_TIMEOUT_MS")) ctx->spin_timeout_ms = atof(t);

  // ---- flatten on the host ----------------------------------------------------------------------
  double phred[128];
  vb2::build_phred_table(phred);
  for (int q = vb2::kNumQual; q < 128; ++q) phred[q] = 1.0;
  vb2::PackedSample &P = ctx->meta;
  std::string perr;
  int rc = vb2::pack_sample(*desc, phred, &P, &perr);
  if (rc) return set_err(ctx, rc, perr);

  // ---- upload ---------------------------------------------------------------------------------
  SampleDev &S = ctx->S;
  const uint32_t *d_words = nullptr, *d_desc = nullptr;
  const double *d_diag = nullptr, *d_kaf = nullptr, *d_phred = nullptr;
  if ((rc = upload(ctx, P.words, &d_words))) return rc;
  if ((rc = upload(ctx, P.slice_desc, &d_desc))) return rc;
  if ((rc = upload(ctx, P.diag, &d_diag))) return rc;
  if ((rc = upload(ctx, P.known_af, &d_kaf))) return rc;
  std::vector<double> phred_v(phred, phred + 128);
  if ((rc = upload(ctx, phred_v, &d_phred, false))) return rc;
  std::vector<float> ud32, mu32;
  const bool fp64 = desc->panel_dtype == VB2_PANEL_FP64;
  if (desc->panel_dtype != VB2_PANEL_FP64 && desc->panel_dtype != VB2_PANEL_FP32)
    return set_err(ctx, VB2_ERR_INVALID, "unknown panel_dtype");
  if (P.known_af.empty()) {
    if (fp64) {
      const double *d_ud = nullptr, *d_mu = nullptr;
      if ((rc = upload(ctx, P.ud, &d_ud))) return rc;
      if ((rc = upload(ctx, P.mu, &d_mu))) return rc;
      S.ud = d_ud; S.mu = d_mu;
    } else {
      ud32.assign(P.ud.begin(), P.ud.end());
      mu32.assign(P.mu.begin(), P.mu.end());
      const float *d_ud = nullptr, *d_mu = nullptr;
      if ((rc = upload(ctx, ud32, &d_ud))) return rc;
      if ((rc = upload(ctx, mu32, &d_mu))) return rc;
      S.ud = d_ud; S.mu = d_mu;
    }
  }
  S.words = d_words;
  S.slice_desc = reinterpret_cast<const uint2 *>(d_desc);
  S.diag = d_diag;
  S.known_af = d_kaf;
  S.phred = d_phred;
  S.log_other_const = P.log_other_const;
  S.min_af = desc->min_af != 0.0 ? desc->min_af : 0.00005;  // h:94
  S.max_af = desc->max_af != 0.0 ? desc->max_af : 0.99995;  // h:95
  S.n_used = P.n_used; S.n_slices = P.n_slices; S.m_pad = P.m_pad; S.n_pc = P.n_pc;
  S.grid_x = (P.n_slices + kWarpsPerCta - 1) / kWarpsPerCta;
  S.panel_fp64 = fp64 ? 1u : 0u;
  uint32_t cap = kStageWordsCap;
  if (const char *t = getenv("VB2_LLK_STAGE_WORDS")) cap = (uint32_t)std::max(1, atoi(t));
  S.stage_words = std::max(1u, std::min(P.max_slice_words, cap));
  S.n_buf = P.max_slice_words > S.stage_words ? 2u : 1u;
  ctx->smem_bytes = kWarpsPerCta * S.n_buf * S.stage_words * 128u;
  VB2_CUDA(ctx, cudaFuncSetAttribute(llk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));

  // ---- result plumbing ---------------------------------------------------------------------------
  VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_mbox, sizeof(Mailbox), cudaHostAllocMapped));
  memset(ctx->h_mbox, 0, sizeof(Mailbox));
  VB2_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->d_mbox, ctx->h_mbox, 0));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_jobs_done, sizeof(unsigned int)));
  VB2_CUDA(ctx, cudaMemsetAsync(ctx->d_jobs_done, 0, sizeof(unsigned int), ctx->stream));
  VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_jobs, sizeof(JobParams) * VB2_MAX_BATCH, cudaHostAllocDefault));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_jobs, sizeof(JobParams) * VB2_MAX_BATCH));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_out, sizeof(double) * VB2_MAX_BATCH));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_sample, sizeof(SampleDev)));
  if ((rc = ensure_slots(ctx, kMaxArgJobs))) return rc;
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // keep only the sizes
  P.words = {}; P.slice_desc = {}; P.ud = {}; P.mu = {}; P.diag = {}; P.known_af = {}; P.marker_index = {};
  return VB2_OK;
}

int vb2_llk_create(const vb2_llk_desc *desc, vb2_llk_ctx **out) {
  if (!out) return set_err(nullptr, VB2_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (!desc) return set_err(nullptr, VB2_ERR_INVALID, "null descriptor");
  if (desc->struct_size != sizeof(vb2_llk_desc))
    return set_err(nullptr, VB2_ERR_INVALID, "vb2_llk_desc.struct_size mismatch (ABI version skew)");
  vb2_llk_ctx *ctx = new (std::nothrow) vb2_llk_ctx();
  if (!ctx) return set_err(nullptr, VB2_ERR_NOMEM, "out of host memory");
  int rc;
  try {
    rc = create_impl(desc, ctx);
  } catch (const std::bad_alloc &) {
    rc = set_err(ctx, VB2_ERR_NOMEM, "out of host memory while flattening the pileup");
  }
  if (rc != VB2_OK) {
    g_last_error = ctx->err;
    vb2_llk_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return VB2_OK;
}

int vb2_llk_get_info(const vb2_llk_ctx *ctx, vb2_llk_info *info) {
  if (!ctx || !info) return set_err(nullptr, VB2_ERR_INVALID, "null argument");
  if (info->struct_size != sizeof(vb2_llk_info)) return set_err(nullptr, VB2_ERR_INVALID, "vb2_llk_info.struct_size mismatch");
  const vb2::PackedSample &P = ctx->meta;
  info->n_pc = P.n_pc;
  info->markers_used = P.n_used;
  info->reads_used = P.reads_used;
  info->reads_streamed = P.reads_streamed;
  info->reads_folded = P.reads_folded;
  info->algorithmic_bytes = 2ull * P.reads_used + 4ull * (P.n_pc + 2) * P.n_used;
  info->device_bytes = ctx->device_bytes;
  info->n_slices = P.n_slices;
  info->grid_x = ctx->S.grid_x;
  info->block_threads = kThreads;
  info->smem_bytes = ctx->smem_bytes;
  info->device = ctx->device;
  info->sm_count = ctx->sm_count;
  info->log_other_const = P.log_other_const;
  return VB2_OK;
}

int vb2_llk_eval_batch(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                       const double *alphas, double *llk_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (!llk_out) return set_err(ctx, VB2_ERR_INVALID, "null output pointer");
  unsigned long long seq = 0;
  int rc = launch_batch(ctx, n, pc_contam, pc_intended, alphas, nullptr, true, &seq);
  if (rc) return rc;
  if (ctx->S.grid_x == 0) {  // no usable marker: the reference's empty sum (h:231, :313)
    for (int j = 0; j < n; ++j) llk_out[j] = 0.0;
    return VB2_OK;
  }
  if ((rc = wait_mailbox(ctx, seq))) return rc;
  for (int j = 0; j < n; ++j) llk_out[j] = ctx->h_mbox->val[j];
  return VB2_OK;
}

int vb2_llk_eval(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha,
                 double *llk_out) {
  return vb2_llk_eval_batch(ctx, 1, pc_contam, pc_intended, &alpha, llk_out);
}

int vb2_llk_eval_batch_device(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                              const double *alphas, double *d_llk_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (!d_llk_out) return set_err(ctx, VB2_ERR_INVALID, "null device output pointer");
  int rc = launch_batch(ctx, n, pc_contam, pc_intended, alphas, d_llk_out, false, nullptr);
  if (rc) return rc;
  if (ctx->S.grid_x == 0) VB2_CUDA(ctx, cudaMemsetAsync(d_llk_out, 0, sizeof(double) * n, ctx->stream));
  return VB2_OK;
}

int vb2_llk_eval_many(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                      const double *alphas, double *llk_out) {
  if (!ctxs || n <= 0 || !ctxs[0]) return set_err(nullptr, VB2_ERR_INVALID, "null/empty context list");
  vb2_llk_ctx *lead = ctxs[0];
  if (n > VB2_MAX_BATCH) return set_err(lead, VB2_ERR_INVALID, "batch size out of range");
  if (!pc_contam || !pc_intended || !alphas || !llk_out) return set_err(lead, VB2_ERR_INVALID, "null argument");
  VB2_CUDA(lead, cudaSetDevice(lead->device));
  if (!lead->h_many) {
    VB2_CUDA(lead, cudaHostAlloc((void **)&lead->h_many, sizeof(SampleDev) * VB2_MAX_BATCH, cudaHostAllocDefault));
    VB2_CUDA(lead, cudaMalloc(&lead->d_many, sizeof(SampleDev) * VB2_MAX_BATCH));
    VB2_CUDA(lead, cudaHostAlloc((void **)&lead->h_slots, sizeof(uint32_t) * VB2_MAX_BATCH, cudaHostAllocDefault));
    VB2_CUDA(lead, cudaMalloc(&lead->d_slots, sizeof(uint32_t) * VB2_MAX_BATCH));
  }
  VB2_CUDA(lead, cudaStreamSynchronize(lead->stream));  // staging buffers are free again
  const uint32_t k = lead->S.n_pc;
  uint32_t grid_x = 0, smem = 0;
  // slot of job j inside its sample = number of earlier jobs on the same context
  for (int j = 0; j < n; ++j) {
    vb2_llk_ctx *c = ctxs[j];
    if (!c) return set_err(lead, VB2_ERR_INVALID, "null context in list");
    if (c->device != lead->device) return set_err(lead, VB2_ERR_INVALID, "contexts live on different devices");
    if (c->S.n_pc != k) return set_err(lead, VB2_ERR_INVALID, "contexts differ in n_pc");
    uint32_t slot = 0;
    for (int i = 0; i < j; ++i) slot += (ctxs[i] == c);
    int rc = ensure_slots(c, slot + 1);
    if (rc) return set_err(lead, rc, c->err);
    lead->h_slots[j] = slot;
  }
  bool any = false;
  for (int j = 0; j < n; ++j) {
    vb2_llk_ctx *c = ctxs[j];
    lead->h_many[j] = c->S;
    // every sample indexes its shared-memory stage with the launch-wide geometry
    grid_x = std::max(grid_x, c->S.grid_x);
    smem = std::max(smem, c->smem_bytes);
    any |= c->S.grid_x > 0;
    fill_job(&lead->h_jobs[j], k, pc_contam + (size_t)j * k, pc_intended + (size_t)j * k, alphas[j]);
  }
  if (!any) {
    for (int j = 0; j < n; ++j) llk_out[j] = 0.0;
    return VB2_OK;
  }
  for (int j = 0; j < n; ++j)
    if (ctxs[j]->S.grid_x == 0)
      return set_err(lead, VB2_ERR_INVALID, "vb2_llk_eval_many: a sample has no usable marker");
  VB2_CUDA(lead, cudaMemcpyAsync(lead->d_many, lead->h_many, sizeof(SampleDev) * n, cudaMemcpyHostToDevice, lead->stream));
  VB2_CUDA(lead, cudaMemcpyAsync(lead->d_slots, lead->h_slots, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, lead->stream));
  VB2_CUDA(lead, cudaMemcpyAsync(lead->d_jobs, lead->h_jobs, sizeof(JobParams) * n, cudaMemcpyHostToDevice, lead->stream));
  LaunchArgs A;
  memset(&A, 0, sizeof(A));
  A.samples = lead->d_many;
  A.slots = lead->d_slots;
  A.jobs_dev = lead->d_jobs;
  A.n_jobs = (uint32_t)n;
  A.mbox = lead->d_mbox;
  A.jobs_done = lead->d_jobs_done;
  A.seq = ++lead->seq;
  dim3 grid(grid_x, (unsigned)n, 1), block(kThreads, 1, 1);
  llk_kernel<<<grid, block, smem, lead->stream>>>(A);
  VB2_CUDA(lead, cudaGetLastError());
  int rc = wait_mailbox(lead, A.seq);
  if (rc) return rc;
  for (int j = 0; j < n; ++j) llk_out[j] = lead->h_mbox->val[j];
  return VB2_OK;
}

int vb2_llk_sync(vb2_llk_ctx *ctx) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VB2_OK;
}

}  // extern "C"
