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

constexpr int kMaxWarps = 4 * vb2::kMaxConcRounds;  // 24 warps: 6 per SM sub-partition
constexpr int kMaxThreads = kMaxWarps * 32;
constexpr int kMaxArgJobs = 4;    // evaluations whose parameters travel in the kernel arguments
constexpr int kMaxArgRounds = 16; // round-table entries that travel in the kernel arguments
constexpr int kNumPairs = 6;      // off-diagonal genotype pairs
constexpr uint32_t kChunkTargetBytes = 4096;  // shared-memory stage per warp

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

struct SampleDev {  // one sample resident in HBM (see llk_pack.h for the blob/round layout)
  const uint8_t *blob;
  const vb2::Round *rounds;  // [n_rounds] in HBM
  double *partials;          // [slots][grid_x]
  unsigned int *tickets;     // [slots]
  double log_other_const, min_af, max_af;
  uint32_t n_rounds, n_bins, grid_x, conc_rounds;
  uint32_t n_pc, panel_fp64, known_af, n_buf;
  uint32_t off_ud, off_mu, off_kaf, off_diag, off_words;
  uint32_t chunk_rows;  // word rows per shared-memory stage
  uint32_t buf_bytes;   // off_words + chunk_rows * 128
  uint32_t pad_;
};
static_assert(sizeof(SampleDev) % 8 == 0 && sizeof(SampleDev) <= 8 * 32, "SampleDev copy loop");

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
  vb2::Round rounds[kMaxArgRounds];  // copy of sample.rounds[0..n_rounds) when it fits
  JobParams jobs[kMaxArgJobs];
  double phred[vb2::kNumQual];       // 10^(-q/10), ContaminationEstimator.h:65-74
};
static_assert(sizeof(LaunchArgs) <= 4000, "kernel arguments must stay below 4 KiB");

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

// acc[i] *= f[i] for the six genotype pairs unless the quality byte is the 0xFF filler:
// one ISETP + six predicated DMULs, no branch.
__device__ __forceinline__ void mul6_if_read(double &a0, double &a1, double &a2, double &a3, double &a4, double &a5,
                                             double f0, double f1, double f2, double f3, double f4, double f5,
                                             uint32_t q) {
  asm("{\n"
      ".reg .pred P;\n"
      "setp.ne.u32 P, %12, 255;\n"
      "@P mul.f64 %0, %0, %6;\n"
      "@P mul.f64 %1, %1, %7;\n"
      "@P mul.f64 %2, %2, %8;\n"
      "@P mul.f64 %3, %3, %9;\n"
      "@P mul.f64 %4, %4, %10;\n"
      "@P mul.f64 %5, %5, %11;\n"
      "}\n"
      : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5)
      : "d"(f0), "d"(f1), "d"(f2), "d"(f3), "d"(f4), "d"(f5), "r"(q));
}

// Four reads (one word) of one lane.  ALT = the word holds alt-allele reads.
template <bool ALT>
__device__ __forceinline__ void eat_word(uint32_t w, const double *s_e, const double (&c0)[kNumPairs],
                                         const double (&c1)[kNumPairs], double (&acc)[kNumPairs]) {
  uint32_t q[4];
  double e[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    q[b] = (w >> (8 * b)) & 0xFFu;
    e[b] = s_e[q[b]];  // s_e[255] = 1.0: the filler byte reads a finite value and is masked below
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    double f[kNumPairs];
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) f[p] = fma(c1[p], e[b], c0[p]);
    if (ALT)
      mul6_if_read(acc[5], acc[4], acc[3], acc[2], acc[1], acc[0], f[0], f[1], f[2], f[3], f[4], f[5], q[b]);
    else
      mul6_if_read(acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], f[0], f[1], f[2], f[3], f[4], f[5], q[b]);
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
__device__ __forceinline__ void marker_af(const uint8_t *blob, const SampleDev &S, const JobParams &J, int lane,
                                          double &af1, double &af2) {
  // h:251-267: AF = (sum_k UD[i][k]*PC[k] + means[i]) / 2, accumulated in k order in fp64.
  const PanelT *ud = reinterpret_cast<const PanelT *>(blob + S.off_ud);
  const PanelT *mu = reinterpret_cast<const PanelT *>(blob + S.off_mu);
  double a1 = 0., a2 = 0.;
  for (uint32_t k = 0; k < S.n_pc; ++k) {
    const double u = (double)ud[k * 32 + lane];
    a1 = __dadd_rn(a1, __dmul_rn(u, J.pc1[k]));
    a2 = __dadd_rn(a2, __dmul_rn(u, J.pc2[k]));
  }
  const double m = (double)mu[lane];
  af1 = (a1 + m) * 0.5;
  af2 = (a2 + m) * 0.5;
}

// One persistent CTA per SM.  Warp w issues on SM sub-partition w % 4 and owns bin
// blockIdx.x*4 + w%4; it serves rounds w/4, w/4 + conc_rounds, ... of that bin.
__global__ void __launch_bounds__(kMaxThreads, 1)
llk_kernel(const __grid_constant__ LaunchArgs A) {
  extern __shared__ __align__(128) uint8_t s_buf[];  // [warp][n_buf][buf_bytes]
  __shared__ double s_e[256];
  __shared__ JobParams s_job;
  __shared__ SampleDev s_sample;
  __shared__ vb2::Round s_rounds[kMaxArgRounds];
  __shared__ double s_red[kMaxWarps];
  __shared__ __align__(8) uint64_t s_bar[kMaxWarps][2];
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5;
  const uint32_t job = blockIdx.y;

  // ---- per-CTA set-up: which sample, which parameters, Phred table ----------------------------
  if (threadIdx.x < sizeof(SampleDev) / 8) {
    const uint64_t *src = A.samples ? reinterpret_cast<const uint64_t *>(A.samples + job)
                                    : reinterpret_cast<const uint64_t *>(&A.sample);
    reinterpret_cast<uint64_t *>(&s_sample)[threadIdx.x] = src[threadIdx.x];
  } else if (threadIdx.x >= 64 && threadIdx.x < 64 + sizeof(JobParams) / 8) {
    const int i = threadIdx.x - 64;
    const double *src = A.jobs_dev ? reinterpret_cast<const double *>(A.jobs_dev + job)
                                   : reinterpret_cast<const double *>(&A.jobs[job < kMaxArgJobs ? job : 0]);
    reinterpret_cast<double *>(&s_job)[i] = src[i];
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_e[i] = i < vb2::kNumQual ? A.phred[i] : 1.0;
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const SampleDev &S = s_sample;
  if (blockIdx.x >= S.grid_x) return;  // eval_many: this sample needs fewer CTAs than the grid has
  const bool rounds_in_smem = S.n_rounds <= kMaxArgRounds;
  if (rounds_in_smem && threadIdx.x < S.n_rounds * (sizeof(vb2::Round) / 8)) {
    const uint64_t *src = A.samples ? reinterpret_cast<const uint64_t *>(S.rounds)
                                    : reinterpret_cast<const uint64_t *>(A.rounds);
    reinterpret_cast<uint64_t *>(s_rounds)[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();

  const uint32_t slot = A.slots ? A.slots[job] : job;
  const uint32_t bin = blockIdx.x * vb2::kBinsPerCta + (warp & 3);
  const uint32_t kc = S.conc_rounds;
  const uint32_t n_rounds = S.n_rounds;
  const uint32_t chunk_rows = S.chunk_rows;
  uint8_t *mybuf = s_buf + (size_t)warp * S.n_buf * S.buf_bytes;

  auto get_round = [&](uint32_t r) -> vb2::Round { return rounds_in_smem ? s_rounds[r] : S.rounds[r]; };
  // first round >= r (stepping by kc) in which this bin owns a blob, or n_rounds
  auto next_round = [&](uint32_t r) -> uint32_t {
    while (r < n_rounds) {
      const vb2::Round R = get_round(r);
      if (bin - R.first_bin < R.count) break;  // unsigned: also false when bin < first_bin
      r += kc;
    }
    return r;
  };
  auto n_chunks = [&](const vb2::Round &R) -> uint32_t {
    return R.rows <= chunk_rows ? 1u : (R.rows + chunk_rows - 1) / chunk_rows;
  };
  // TMA: chunk c of this bin's blob of round r -> buffer b.  Chunk 0 = header + panel + diag + the
  // first chunk_rows word rows; chunk c >= 1 = the next chunk_rows word rows.
  auto issue = [&](uint32_t r, uint32_t c, uint32_t b) {
    const vb2::Round R = get_round(r);
    const uint8_t *src = S.blob + R.base + (uint64_t)(bin - R.first_bin) * R.stride;
    uint32_t off, bytes;
    if (c == 0) {
      off = 0;
      bytes = S.off_words + (R.rows < chunk_rows ? R.rows : chunk_rows) * 128u;
    } else {
      off = S.off_words + c * chunk_rows * 128u;
      const uint32_t left = R.rows - c * chunk_rows;
      bytes = (left < chunk_rows ? left : chunk_rows) * 128u;
    }
    mbar_arrive_expect_tx(&s_bar[warp][b], bytes);
    bulk_g2s(mybuf + (size_t)b * S.buf_bytes, src + off, bytes, &s_bar[warp][b]);
  };

  double vsum = 0.0;
  if ((uint32_t)(warp >> 2) < kc) {
    double c0[kNumPairs], c1[kNumPairs];
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) {
      c0[p] = s_job.c0[p];
      c1[p] = s_job.c1[p];
    }
    // consume cursor (r, c) and issue cursor (ir, ic); both advance through the same sequence
    uint32_t r = next_round((uint32_t)(warp >> 2)), c = 0;
    uint32_t ir = r, ic = 0, ib = 0, cb = 0, parity = 0;
    auto advance_issue = [&]() {
      if (ir >= n_rounds) return;
      const vb2::Round R = get_round(ir);
      if (ic + 1 < n_chunks(R)) { ++ic; } else { ir = next_round(ir + kc); ic = 0; }
    };
    if (ir < n_rounds) {
      if (lane == 0) issue(ir, ic, ib);
      advance_issue();
      ib ^= (S.n_buf - 1);
    }
    double acc[kNumPairs], af1 = 0., af2 = 0., ldiag = 0.;
    uint32_t wr = 0, wa = 0, n_valid = 0;
    while (r < n_rounds) {
      if (ir < n_rounds && S.n_buf == 2) {
        __syncwarp();  // every lane finished reading the buffer about to be overwritten
        if (lane == 0) issue(ir, ic, ib);
        advance_issue();
        ib ^= 1u;
      }
      mbar_wait(&s_bar[warp][cb], (parity >> cb) & 1u);
      parity ^= 1u << cb;
      const uint8_t *buf = mybuf + (size_t)cb * S.buf_bytes;
      const vb2::Round R = get_round(r);
      if (c == 0) {
        // ---- (i) header, allele frequencies, genotype priors, diagonal pairs ----------------------
        const uint4 hdr = *reinterpret_cast<const uint4 *>(buf);
        wr = hdr.x; wa = hdr.y; n_valid = hdr.z;
        if (S.known_af) {
          af1 = af2 = reinterpret_cast<const double *>(buf + S.off_kaf)[lane];  // h:251-252
        } else if (S.panel_fp64) {
          marker_af<double>(buf, S, s_job, lane, af1, af2);
        } else {
          marker_af<float>(buf, S, s_job, lane, af1, af2);
        }
        double gf[3], gf2[3];
        initial_gf(af1, S.min_af, S.max_af, gf);   // contaminating sample
        initial_gf(af2, S.min_af, S.max_af, gf2);  // intended sample
        const double *dg = reinterpret_cast<const double *>(buf + S.off_diag);
        ldiag = dg[lane] * (gf[0] * gf2[0]) + dg[32 + lane] * (gf[1] * gf2[1]) + dg[64 + lane] * (gf[2] * gf2[2]);
#pragma unroll
        for (int p = 0; p < kNumPairs; ++p) acc[p] = 1.0;
      }
      // ---- (ii) stream this chunk's word rows: ref rows first, then alt rows --------------------
      const uint32_t t_lo = c * chunk_rows;
      uint32_t t_hi = t_lo + chunk_rows;
      if (t_hi > wr + wa) t_hi = wr + wa;
      const uint32_t *rows = reinterpret_cast<const uint32_t *>(buf + (c == 0 ? S.off_words : 0u)) + lane;
      const uint32_t t_ref_hi = wr < t_hi ? wr : t_hi;
      uint32_t t = t_lo;
      for (; t < t_ref_hi; ++t) eat_word<false>(rows[(t - t_lo) * 32], s_e, c0, c1, acc);
      for (; t < t_hi; ++t) eat_word<true>(rows[(t - t_lo) * 32], s_e, c0, c1, acc);

      if (c + 1 == n_chunks(R)) {
        // ---- (iii) marginal over the nine genotype pairs, log -------------------------------------
        // h:307-311: markerLK = sum_{g1,g2} exp(acc) * GF[g1] * GF2[g2]; here exp(acc) is the running
        // product itself, and the diagonal pairs were folded into ldiag above.
        double gf[3], gf2[3];
        initial_gf(af1, S.min_af, S.max_af, gf);
        initial_gf(af2, S.min_af, S.max_af, gf2);
        double L = ldiag;
#pragma unroll
        for (int p = 0; p < kNumPairs; ++p) L += acc[p] * (gf[pair_g1(p)] * gf2[pair_g2(p)]);
        if ((uint32_t)lane < n_valid && L > 0) vsum += log(L);
        r = next_round(r + kc);
        c = 0;
      } else {
        ++c;
      }
      if (S.n_buf == 2) cb ^= 1u;
    }
  }

  // ---- fixed-order reduction: warp shuffle tree -> CTA -> last CTA sums the grid ---------------
#pragma unroll
  for (int o = 16; o; o >>= 1) vsum += __shfl_xor_sync(0xFFFFFFFFu, vsum, o);
  if (lane == 0) s_red[warp] = vsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double cta = 0.0;
    for (int w = 0; w < n_warps; ++w) cta += s_red[w];
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
  uint32_t block_threads = 0;
  std::vector<vb2::Round> rounds;  // host copy (kernel arguments)
  double phred[vb2::kNumQual];
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
  memcpy(A.phred, ctx->phred, sizeof(A.phred));
  if (ctx->rounds.size() <= (size_t)kMaxArgRounds)
    memcpy(A.rounds, ctx->rounds.data(), ctx->rounds.size() * sizeof(vb2::Round));
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
  dim3 grid(ctx->S.grid_x, (unsigned)n, 1), block(ctx->block_threads, 1, 1);
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
  vb2::build_phred_table(ctx->phred);
  if (desc->panel_dtype != VB2_PANEL_FP64 && desc->panel_dtype != VB2_PANEL_FP32)
    return set_err(ctx, VB2_ERR_INVALID, "unknown panel_dtype");
  vb2::PackConfig cfg;
  cfg.max_ctas = (uint32_t)ctx->sm_count;  // one persistent CTA per SM
  cfg.panel_fp64 = desc->panel_dtype == VB2_PANEL_FP64;
  if (const char *t = getenv("VB2_LLK_MAX_CTAS")) cfg.max_ctas = (uint32_t)std::max(1, atoi(t));
  vb2::PackedSample &P = ctx->meta;
  std::string perr;
  int rc = vb2::pack_sample(*desc, cfg, ctx->phred, &P, &perr);
  if (rc) return set_err(ctx, rc, perr);
  ctx->rounds = P.rounds;

  // ---- upload ---------------------------------------------------------------------------------
  SampleDev &S = ctx->S;
  const uint8_t *d_blob = nullptr;
  const vb2::Round *d_rounds = nullptr;
  if ((rc = upload(ctx, P.blob, &d_blob))) return rc;
  if ((rc = upload(ctx, P.rounds, &d_rounds, false))) return rc;
  S.blob = d_blob;
  S.rounds = d_rounds;
  S.log_other_const = P.log_other_const;
  S.min_af = desc->min_af != 0.0 ? desc->min_af : 0.00005;  // h:94
  S.max_af = desc->max_af != 0.0 ? desc->max_af : 0.99995;  // h:95
  S.n_rounds = (uint32_t)P.rounds.size();
  S.n_bins = P.n_bins;
  S.grid_x = P.n_slices ? P.grid_x : 0u;
  S.conc_rounds = P.conc_rounds;
  S.n_pc = P.n_pc;
  S.panel_fp64 = cfg.panel_fp64 ? 1u : 0u;
  S.known_af = P.known_af ? 1u : 0u;
  S.off_ud = P.layout.off_ud; S.off_mu = P.layout.off_mu; S.off_kaf = P.layout.off_kaf;
  S.off_diag = P.layout.off_diag; S.off_words = P.layout.off_words;
  // shared-memory stage: as many word rows as fit next to the fixed part in ~4 KiB (at least 4),
  // never more than the deepest blob needs; VB2_LLK_STAGE_WORDS overrides (tests)
  uint32_t max_rows = 0;
  for (const vb2::Round &R : P.rounds) max_rows = std::max(max_rows, R.rows);
  uint32_t cap_rows = kChunkTargetBytes > S.off_words + 4u * 128u ? (kChunkTargetBytes - S.off_words) / 128u : 4u;
  if (const char *t = getenv("VB2_LLK_STAGE_WORDS")) cap_rows = (uint32_t)std::max(1, atoi(t));
  S.chunk_rows = std::max(1u, std::min(std::max(max_rows, 1u), cap_rows));
  S.buf_bytes = S.off_words + S.chunk_rows * 128u;
  const bool one_item_per_warp = S.n_rounds <= S.conc_rounds && max_rows <= S.chunk_rows;
  S.n_buf = one_item_per_warp ? 1u : 2u;
  ctx->block_threads = 128u * std::max(1u, S.conc_rounds);
  ctx->smem_bytes = (ctx->block_threads / 32u) * S.n_buf * S.buf_bytes;
  if (ctx->smem_bytes > 200u * 1024u) return set_err(ctx, VB2_ERR_INVALID, "shared-memory stage too large (n_pc too big?)");
  VB2_CUDA(ctx, cudaFuncSetAttribute(llk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

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
  P.blob = {}; P.marker_index = {};
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
  info->block_threads = ctx->block_threads;
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
  uint32_t grid_x = 0, smem = 0, threads = 0;
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
    threads = std::max(threads, c->block_threads);
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
  memcpy(A.phred, lead->phred, sizeof(A.phred));
  dim3 grid(grid_x, (unsigned)n, 1), block(threads, 1, 1);
  llk_kernel<<<grid, block, smem, lead->stream>>>(A);
  VB2_CUDA(lead, cudaGetLastError());
  int rc = wait_mailbox(lead, A.seq);
  if (rc) return rc;
  for (int j = 0; j < n; ++j) llk_out[j] = lead->h_mbox->val[j];
  return VB2_OK;
}


int vb2_llk_time_device(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup, int steps, const double *pc_contam,
                        const double *pc_intended, double alpha, float *elapsed_ms) {
  if (!ctxs || n_ctx <= 0 || !ctxs[0] || steps <= 0 || !elapsed_ms) return set_err(nullptr, VB2_ERR_INVALID, "bad argument");
  vb2_llk_ctx *lead = ctxs[0];
  VB2_CUDA(lead, cudaSetDevice(lead->device));
  for (int i = 0; i < n_ctx; ++i)
    if (!ctxs[i] || ctxs[i]->stream != lead->stream)
      return set_err(lead, VB2_ERR_INVALID, "vb2_llk_time_device: contexts must share one stream");
  cudaEvent_t e0, e1;
  VB2_CUDA(lead, cudaEventCreate(&e0));
  VB2_CUDA(lead, cudaEventCreate(&e1));
  std::vector<double> pc1(pc_contam, pc_contam + lead->S.n_pc);
  int rc = VB2_OK;
  for (int i = -warmup; i < steps && rc == VB2_OK; ++i) {
    if (i == 0) cudaEventRecord(e0, lead->stream);
    pc1[0] = pc_contam[0] + 1e-7 * ((i + warmup) % 1000);  // a different point every step
    vb2_llk_ctx *c = ctxs[(i + warmup) % n_ctx];
    rc = launch_batch(c, 1, pc1.data(), pc_intended, &alpha, c->d_out, false, nullptr);
  }
  cudaEventRecord(e1, lead->stream);
  cudaError_t e = cudaEventSynchronize(e1);
  if (rc == VB2_OK && e != cudaSuccess) rc = set_err(lead, VB2_ERR_CUDA, cudaGetErrorString(e));
  if (rc == VB2_OK) cudaEventElapsedTime(elapsed_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int vb2_llk_time_host(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup, int steps, const double *pc_contam,
                      const double *pc_intended, double alpha, double *elapsed_s, double *last_llk) {
  if (!ctxs || n_ctx <= 0 || !ctxs[0] || steps <= 0 || !elapsed_s) return set_err(nullptr, VB2_ERR_INVALID, "bad argument");
  vb2_llk_ctx *lead = ctxs[0];
  std::vector<double> pc1(pc_contam, pc_contam + lead->S.n_pc);
  double llk = 0.0;
  std::chrono::steady_clock::time_point t0;
  for (int i = -warmup; i < steps; ++i) {
    if (i == 0) {
      VB2_CUDA(lead, cudaStreamSynchronize(lead->stream));
      t0 = std::chrono::steady_clock::now();
    }
    pc1[0] = pc_contam[0] + 1e-7 * ((i + warmup) % 1000);
    int rc = vb2_llk_eval(ctxs[(i + warmup) % n_ctx], pc1.data(), pc_intended, alpha, &llk);
    if (rc) return rc;
  }
  *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (last_llk) *last_llk = llk;
  return VB2_OK;
}

int vb2_llk_sync(vb2_llk_ctx *ctx) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VB2_OK;
}

}  // extern "C"
