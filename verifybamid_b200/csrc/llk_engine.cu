// llk_engine.cu -- the sm_100a contamination-likelihood kernels and the C ABI around them
// (include/vb2_llk.h).  Replaces, for one sample resident in HBM,
//     FullLLKFunc::ComputeMixLLKs            reference ContaminationEstimator.h:194-314
// What every kernel computes for a 32-marker slice (one warp, one marker per lane):
//   (i)   AF = (UD.PC + mu)/2 per marker (h:251-267) and Hardy-Weinberg priors (h:186-192);
//   (ii)  per read, the six alpha-dependent genotype-pair emissions of the 3x3 mixture
//         (h:213-229, in the closed form of SURVEY.md Appendix A: each is LINEAR in the Phred
//         error e, F_p(e) = c0_p + c1_p*e; two reads are eaten at once through the symmetric
//         functions of their errors: 10 fp64 instructions per read for all six pairs);
//         everything a warp needs for 32 markers is one contiguous blob fetched by ONE TMA
//         bulk copy (cp.async.bulk + mbarrier) whose address is pure arithmetic;
//   (iii) the marginal per marker (h:307-311); the marginals of a bin are multiplied up and one log
//         per lane is taken; fixed-order lane / bin / CTA reduction in fp64 (h:232-236 is an OpenMP
//         reduction), on the device or by the host over a host-mapped mailbox.
// Four launch shapes of the same arithmetic (identical bits):
//   llk_kernel          one evaluation per launch, one CTA per SM;
//   llk_flow_kernel     many evaluations per launch, the common shapes (fp32 panel, NumPC 2 or 4, blobs that fit a
//                       stage): per-job coefficients in the kernel arguments -> uniform registers, 64 registers,
//                       eight warps per SM sub-partition, launches of a batch overlapped;
//   llk_stream_kernel   many evaluations per launch, any shape: persistent grid, warps pull (evaluation, bin) tasks
//                       from a queue, task records by TMA; llk_reduce_kernel adds the per-bin partial sums of both;
//   llk_session_kernel  resident: the sample stays in shared memory, evaluations arrive through a
//                       host-mapped doorbell (vb2_llk_session_begin / _end), or a whole Nelder-Mead search runs
//                       next to it (vb2_llk_minimize).
// What bounds them on a B200 is FP64 instruction issue (an FP64 warp-instruction holds the issue slot for two cycles,
// three with three register operands: tools/microbench_fp64.cu, microbench_mix.cu).
// Everything that is evaluation-invariant was folded at create time by llk_pack.cpp.
//
// There is NO CPU fallback in this file: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "llk_internal.h"
#include "llk_pack.h"
#include "vb2_llk.h"

namespace {

constexpr int kMaxConcRounds = (int)vb2::kMaxConcRounds;  // rounds of a bin a one-evaluation CTA keeps in flight
constexpr int kMaxWarps = 4 * kMaxConcRounds;  // 16 warps (4 per SM sub-partition): 128 registers per thread
constexpr int kMaxThreads = kMaxWarps * 32;
#ifndef VB2_SESSION_MAX_CONC
#define VB2_SESSION_MAX_CONC 4   // rounds of a bin the resident kernel runs concurrently (4 warps each)
#endif
constexpr int kSessionMaxConc = VB2_SESSION_MAX_CONC;
constexpr int kSessionMaxWarps = 4 * kSessionMaxConc;
constexpr int kMaxArgJobs = 4;    // evaluations whose parameters travel in the kernel arguments
constexpr int kMaxArgRounds = 16; // round-table entries that travel in the kernel arguments
constexpr int kNumPairs = 6;      // off-diagonal genotype pairs
constexpr uint32_t kChunkTargetBytes = 4096;  // shared-memory stage per warp
constexpr int kTraceSlots = 16;
constexpr uint32_t kQueueWords = 64;  // task counters of a many-evaluations batch (one per launch it is cut into)
constexpr int kPhredArgs = 96;  // Phred errors 0..93 (+2 pad) that travel in the kernel arguments
constexpr double kLn2 = 0.693147180559945309417232121458;

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
  double *partials;          // [slots][4*grid_x] (device-side reduction only; llk_kernel uses [slots][grid_x] of it)
  unsigned int *tickets;     // [slots]
  double log_other_const, min_af, max_af;
  uint32_t n_rounds, n_bins, grid_x, conc_rounds;
  uint32_t n_pc, panel_fp64, known_af, n_buf;
  uint32_t off_ud, off_mu, off_kaf, off_diag, off_words;
  uint32_t chunk_rows;  // word rows per shared-memory stage
  uint32_t buf_bytes;   // off_words + (chunk_rows + 1) * 128
  uint32_t pad_;
};
static_assert(sizeof(SampleDev) % 8 == 0 && sizeof(SampleDev) <= 8 * 32, "SampleDev copy loop");

// Host-mapped result slot.  Value and sequence number travel in ONE 16-byte store, so the host,
// which polls `seq`, never sees a new sequence number next to an old value.
struct __align__(16) Slot {
  double val;
  unsigned long long seq;
};

// One job of a many-evaluations launch, as llk_stream_kernel prefetches it: everything a warp needs to know to run
// a task of this job sits in ONE 16-byte aligned record that a single TMA bulk copy brings into shared memory a
// whole task ahead of its use -- no dependent chain of global loads (sample table -> round table -> parameters)
// at a task switch.  Round tables longer than kMaxArgRounds continue in HBM (S.rounds).
constexpr int kRecRounds = 16;
struct __align__(16) TaskRec {
  SampleDev S;                      // the job's sample
  uint32_t pslot, pad_;             // its partial-sum slot inside the sample
  vb2::Round rounds[kRecRounds];    // S.rounds[0 .. min(n_rounds, kRecRounds))
  JobParams J;                      // the job's parameters
};
static_assert(sizeof(TaskRec) % 16 == 0, "TaskRec is copied with cp.async.bulk");

// Marker shards on several GPUs of one box (one process per GPU): instead of an NCCL all-reduce behind the kernel, the
// reduce kernel itself PUSHES every job's shard sum into every rank's buffer over NVLink (peer stores), stamps a flag
// when all of a launch's sums are out, and a small gather kernel on every rank adds the shards in rank order --
// a one-shot all-gather + local sum, two banks deep.  Buffers are exchanged as CUDA IPC handles (vb2_peer_*).
constexpr int kMaxPeers = 8;
struct PeerDev {
  double *vals[kMaxPeers];               // rank p's buffer: [2 banks][world][VB2_MAX_BATCH] shard sums ...
  unsigned long long *flags[kMaxPeers];  // ... and [2 banks][world] launch stamps behind them
  unsigned int *ticket;                  // (local) CTAs of the reduce kernel that have pushed their job
  uint32_t world, rank;
};

struct LaunchArgs {
  SampleDev sample;           // ARGS kernels: the sample itself
  const SampleDev *samples;   // generic llk_kernel: job j evaluates samples[j]
  const uint32_t *slots;      // generic llk_kernel: partial/ticket slot of job j inside its sample
  const JobParams *jobs_dev;  // generic llk_kernel: parameters in HBM
  const TaskRec *recs;        // llk_stream_kernel / llk_reduce_kernel: job j = recs[j]
  const PeerDev *peer;        // llk_reduce_kernel: push the shard sums to the peers (else nullptr)
  unsigned long long peer_seq;  // the stamp of this launch on the peers' flags (its parity picks the bank)
  double *d_out;              // [n_jobs] device results (device-side reduction; may be nullptr)
  Slot *mbox;                 // device view of the host mailbox (may be nullptr)
  unsigned long long seq;
  uint32_t n_jobs;
  uint32_t kc;     // rounds a CTA runs concurrently in THIS launch (it has 4*kc warps)
  uint32_t n_buf;  // shared-memory stages per warp in THIS launch (1 or 2)
  uint32_t n_bins_max;   // llk_stream_kernel: bins per evaluation in this launch (tasks = n_jobs * n_bins_max)
  uint32_t stage_bytes;  // llk_stream_kernel: bytes per shared-memory stage (largest buf_bytes of the launch)
  uint32_t pad_;
  unsigned long long *trace;  // llk_kernel: [grid_x][kTraceSlots] clock stamps of thread 0 (diagnostics; usually null)
  unsigned int *queue;   // llk_stream_kernel: next task to hand out; zero between launches (llk_reduce_kernel rewinds it)
  double phred[kPhredArgs];          // 10^(-q/10), q = 0..93, exactly as the host computes it (h:65-74)
  vb2::Round rounds[kMaxArgRounds];  // ARGS kernels: sample.rounds[0..n_rounds)
  JobParams jobs[kMaxArgJobs];       // ARGS kernels: parameters of job blockIdx.y
};
static_assert(sizeof(LaunchArgs) <= 4000, "kernel arguments must stay below 4 KiB");
static_assert(kRecRounds == kMaxArgRounds, "one round-table size");

// 10^(-q/10), q = 0..93, travels in the kernel arguments (LaunchArgs::phred) exactly as the host computes it
// (ContaminationEstimator.h:65-74); entries 94..255 of the shared-memory copy are 1.0 (a 0xFF filler byte indexes 255).
// The per-CTA copy every kernel reads: at file scope, so that a look-up is ONE instruction,
// LDS.64 [byte offset register + constant].
__shared__ double s_e[256];

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
// one non-blocking look at an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Wait with a warp-uniform exit: every lane polls (try_wait suspends the thread for a hardware time slice) and the
// warp leaves together on a vote -- so the compiler knows the warp is converged behind the wait.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {  // bar = shared-window address
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (__all_sync(0xFFFFFFFFu, ok != 0)) break;
  }
}
// expect-tx + bulk copy as ONE predicated pair (no branch: the caller's basic block stays whole)
__device__ __forceinline__ void bulk_g2s_if(bool pred, void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "setp.ne.u32 P1, %4, 0;\n"
      "@P1 mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
      "@P1 cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "r"((uint32_t)pred) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// (ii) the read stream
// ---------------------------------------------------------------------------------------------
// For a ref-class read with Phred error e the emission of genotype pair p is F_p(e) = c0_p + c1_p e.
// Two reads are eaten at once through the symmetric functions of their errors,
//     F_p(ea) F_p(eb) = C0_p + C1_p (ea + eb) + C2_p ea eb,     C0 = c0^2, C1 = c0 c1, C2 = c1^2,
// so a word of four reads costs 4 shared instructions (two sums, two products) plus, per pair,
// 4 DFMA + 2 DMUL: 40 fp64 instructions instead of the 48 of four separate F_p(e) factors.
// log of a (sub)normal marginal: never on the hot path, kept out of line so its code is not fetched
__device__ __noinline__ double cold_log(double x) { return log(x); }

struct Quad {
  static constexpr bool kPrefetchRows = true;  // the read loop fetches row t+1 before the arithmetic of row t
  static constexpr bool kDotAddress = true;    // Phred look-up address by IDP.4A (see phred_of)
  static constexpr uint32_t sel[4] = {0x8u, 0x800u, 0x80000u, 0x8000000u};  // (only the plain loop reads them)
  double C0[kNumPairs], C1[kNumPairs], C2[kNumPairs];
};
struct StreamQuad : Quad {                     // llk_stream_kernel's holder
  static constexpr bool kDotAddress = false;
};

// Four reads of one lane.  ALT = alt-allele reads: acc[5-p] takes what a ref read gives acc[p].
// Written breadth-first (all pairs advance together) so that dependent instructions sit >= 6 apart.
template <bool ALT, typename QT>
__device__ __forceinline__ void eat4(double e0, double e1, double e2, double e3, const QT &Q,
                                     double (&acc)[kNumPairs]) {
  const double s01 = e0 + e1, t01 = e0 * e1, s23 = e2 + e3, t23 = e2 * e3;
  double g[kNumPairs], h[kNumPairs];
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p) {
    g[p] = fma(Q.C1[p], s01, Q.C0[p]);
    h[p] = fma(Q.C1[p], s23, Q.C0[p]);
  }
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p) {
    g[p] = fma(Q.C2[p], t01, g[p]);
    h[p] = fma(Q.C2[p], t23, h[p]);
  }
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p) g[p] *= h[p];
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p) acc[ALT ? (kNumPairs - 1 - p) : p] *= g[p];
}

// n rows in which every lane holds four real reads; `col` = this lane's column of the first row.  The
// word and the four table look-ups of row t+1 are issued before the arithmetic of row t -- also after the last
// row: the loop reads one row past the run (the next run, or the 128-byte pad every stage buffer ends with) and
// drops what it read, which keeps the loop free of a peeled copy.
// Phred error of byte b (0..3) of word w.
// DOT: measured per kernel (tools/gpu_ab_prmt.sh) -- the byte dot product is faster in llk_flow_kernel and in the
// resident kernel (7.60 vs 7.93 us per dependent evaluation), the extract + multiply-add pair in llk_stream_kernel
// (3.19-3.24 vs 3.28 us), so the coefficient holder of a kernel says which one its read loop uses (kDotAddress).
template <int B, bool DOT = true>
__device__ __forceinline__ double phred_of(uint32_t w) {
  uint32_t off;
  if (DOT) {
    // 8 * byte B of w in ONE instruction: a four-way byte dot product with (8 in position B, 0 elsewhere) -- IDP.4A
    off = (uint32_t)__dp4a(w, 0x8u << (8 * B), 0u);
  } else {
    off = __byte_perm(w, 0u, 0x4440u + B) * 8u;  // (0, 0, 0, byte B) * 8
  }
  return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(s_e) + off);
}

// the same look-up with the selector (8 << 8*B) in a register
__device__ __forceinline__ double phred_at(uint32_t w, uint32_t sel) {
  return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(s_e) + (uint32_t)__dp4a(w, sel, 0u));
}
#ifndef VB2_FLOW_LOOP_UNROLL
#define VB2_FLOW_LOOP_UNROLL 1
#endif
constexpr int kFlowLoopUnroll = VB2_FLOW_LOOP_UNROLL;  // rows per iteration of the plain read loop (A/B builds)
template <bool ALT, typename QT>
__device__ __forceinline__ void eat_full_rows(const uint32_t *col, uint32_t n, const QT &Q,
                                              double (&acc)[kNumPairs]) {
  if constexpr (!QT::kPrefetchRows) {  // (enough warps per scheduler to cover the look-ups: no row in flight, fewer registers)
#pragma unroll kFlowLoopUnroll
    for (uint32_t t = 0; t < n; ++t) {
      const uint32_t w = col[t * 32];
      // (the four dot-product selectors are loop-invariant REGISTERS of the holder: as immediates ptxas re-creates them in
      //  uniform registers every iteration -- four extra instructions per row)
      eat4<ALT>(phred_at(w, Q.sel[0]), phred_at(w, Q.sel[1]), phred_at(w, Q.sel[2]), phred_at(w, Q.sel[3]), Q, acc);
    }
    return;
  }
  if (n == 0) return;
  uint32_t w = col[0];
  double e0 = phred_of<0, QT::kDotAddress>(w), e1 = phred_of<1, QT::kDotAddress>(w), e2 = phred_of<2, QT::kDotAddress>(w), e3 = phred_of<3, QT::kDotAddress>(w);
#pragma unroll 1
  for (uint32_t t = 1; t <= n; ++t) {
    w = col[t * 32];
    const double n0 = phred_of<0, QT::kDotAddress>(w), n1 = phred_of<1, QT::kDotAddress>(w), n2 = phred_of<2, QT::kDotAddress>(w),
                 n3 = phred_of<3, QT::kDotAddress>(w);
    eat4<ALT>(e0, e1, e2, e3, Q, acc);
    e0 = n0; e1 = n1; e2 = n2; e3 = n3;
  }
}

// The ragged last word of a run when every lane holds the same number n (1..3) of reads in it: the
// first n bytes are reads in all lanes, so no byte has to be inspected.  lin = {c0[6], c1[6]} (shared memory).
template <bool ALT, typename QT>
__device__ __forceinline__ void eat_word_tail(uint32_t w, uint32_t n, const double *lin, const QT &Q,
                                              double (&acc)[kNumPairs]) {
  const double e0 = s_e[w & 0xFFu];
  if (n == 1) {
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) acc[ALT ? (kNumPairs - 1 - p) : p] *= fma(lin[kNumPairs + p], e0, lin[p]);
    return;
  }
  const double e1 = s_e[(w >> 8) & 0xFFu];
  const double s = e0 + e1, t = e0 * e1;
  if (n == 2) {
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p)
      acc[ALT ? (kNumPairs - 1 - p) : p] *= fma(Q.C2[p], t, fma(Q.C1[p], s, Q.C0[p]));
    return;
  }
  const double e2 = s_e[(w >> 16) & 0xFFu];
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p)
    acc[ALT ? (kNumPairs - 1 - p) : p] *= fma(Q.C2[p], t, fma(Q.C1[p], s, Q.C0[p])) * fma(lin[kNumPairs + p], e2, lin[p]);
}

// Rows that may carry 0xFF filler bytes (the last words of a lane whose run is shorter than its slice's).
// Branch-free: every byte gives a factor F_p(e) = c0_p + c1_p e, replaced by 1 where the byte is a filler, so
// the lanes of a warp never diverge (the slices with such rows are few but sit together in a handful of bins,
// whose CTAs would otherwise finish a quarter later than the rest).
template <bool ALT>
__device__ __forceinline__ void eat_checked(const uint32_t *col, uint32_t n_rows, const double *lin,
                                            double (&acc)[kNumPairs]) {
#pragma unroll 1
  for (uint32_t t = 0; t < n_rows; ++t) {
    const uint32_t w = col[t * 32];
    const uint32_t q0 = w & 0xFFu, q1 = (w >> 8) & 0xFFu, q2 = (w >> 16) & 0xFFu, q3 = w >> 24;
    const double e0 = s_e[q0], e1 = s_e[q1], e2 = s_e[q2], e3 = s_e[q3];
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) {
      const double c0 = lin[p], c1 = lin[kNumPairs + p];
      const double f0 = q0 != 0xFFu ? fma(c1, e0, c0) : 1.0, f1 = q1 != 0xFFu ? fma(c1, e1, c0) : 1.0;
      const double f2 = q2 != 0xFFu ? fma(c1, e2, c0) : 1.0, f3 = q3 != 0xFFu ? fma(c1, e3, c0) : 1.0;
      acc[ALT ? (kNumPairs - 1 - p) : p] *= (f0 * f1) * (f2 * f3);
    }
  }
}

// One run (the ref-allele or the alt-allele reads of the slice): n_full rows in which every lane holds four real
// reads, then n_ragged rows that may hold fillers; when `tail` is 1..3 the (single) ragged row holds exactly
// that many reads in every lane.
template <bool ALT, typename QT>
__device__ __forceinline__ void eat_rows(const uint32_t *col, uint32_t n_full, uint32_t n_ragged, uint32_t tail,
                                         const double *lin, const QT &Q, double (&acc)[kNumPairs]) {
  eat_full_rows<ALT>(col, n_full, Q, acc);
  if (n_ragged == 0) return;
  col += (size_t)n_full * 32;
  if (tail && n_ragged == 1) eat_word_tail<ALT>(col[0], tail, lin, Q, acc);
  else eat_checked<ALT>(col, n_ragged, lin, acc);
}

// Both runs of a slice.  MERGED = false: two copies of the code (ref, alt) with the accumulators renamed at
// compile time.  MERGED = true: ONE copy run twice with the accumulators reversed in between (acc[p] <-> acc[5-p]
// is exactly what turns a ref read into an alt read) -- half the instructions to fetch, which is what bounds a
// launch that evaluates once: its code arrives cold from L2 at about five cycles per instruction.
template <bool MERGED, typename QT>
__device__ __forceinline__ void eat_runs(const uint32_t *col, uint32_t fr, uint32_t rr, uint32_t tr, uint32_t fa,
                                         uint32_t ra, uint32_t ta, const double *lin, const QT &Q,
                                         double (&acc)[kNumPairs]) {
  if constexpr (!MERGED) {
    eat_rows<false>(col, fr, rr, tr, lin, Q, acc);
    eat_rows<true>(col + (size_t)(fr + rr) * 32, fa, ra, ta, lin, Q, acc);
  } else {
#pragma unroll 1
    for (int cls = 0; cls < 2; ++cls) {
      eat_rows<false>(col, fr, rr, tr, lin, Q, acc);
      col += (size_t)(fr + rr) * 32;
      fr = fa; rr = ra; tr = ta;
      double t;
      t = acc[0]; acc[0] = acc[5]; acc[5] = t;
      t = acc[1]; acc[1] = acc[4]; acc[4] = t;
      t = acc[2]; acc[2] = acc[3]; acc[3] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (i) per-marker set-up
// ---------------------------------------------------------------------------------------------
// ContaminationEstimator.h:186-192 with the reference's comparison order (NaN passes through).
__device__ __forceinline__ void initial_gf(double af, double min_af, double max_af, double (&gf)[3]) {
  if (af < min_af) af = min_af;
  if (af > max_af) af = max_af;
  gf[0] = __dmul_rn(1 - af, 1 - af);
  gf[1] = __dmul_rn(__dmul_rn(2, af), 1 - af);
  gf[2] = __dmul_rn(af, af);
}

// Blob layout as the kernel sees it.  FixedLayout<NPC>: the common case (fp32 panel, default AF clamps,
// no --KnownAF, NumPC = NPC) with every offset a compile-time constant and the PC loop unrolled;
// RuntimeLayout: anything else, loaded once per warp from the sample descriptor.
template <int NPC>
struct FixedLayout {
  static constexpr bool kFixed = true;
  static constexpr uint32_t n_pc = NPC, off_ud = vb2::kBlobHeaderBytes, off_mu = off_ud + NPC * 128u,
                            off_diag = off_mu + 128u, off_words = off_diag + 768u, off_kaf = 0u;
  static constexpr bool panel_fp64 = false, known_af = false;
  static constexpr double min_af = 0.00005, max_af = 0.99995;  // h:94-95
  __device__ __forceinline__ explicit FixedLayout(const SampleDev &) {}
};
struct RuntimeLayout {
  static constexpr bool kFixed = false;
  uint32_t n_pc, off_ud, off_mu, off_kaf, off_diag, off_words;
  bool panel_fp64, known_af;
  double min_af, max_af;
  __device__ __forceinline__ explicit RuntimeLayout(const SampleDev &S)
      : n_pc(S.n_pc), off_ud(S.off_ud), off_mu(S.off_mu), off_kaf(S.off_kaf), off_diag(S.off_diag),
        off_words(S.off_words), panel_fp64(S.panel_fp64 != 0), known_af(S.known_af != 0), min_af(S.min_af),
        max_af(S.max_af) {}
};

// h:251-267: AF = (sum_k UD[i][k]*PC[k] + means[i]) / 2.
//   fp64 panel: the reference's operation order exactly (products and sums rounded separately, k ascending,
//               the mean added last), so AF is the reference's bit for bit;
//   fp32 panel: the stored panel is already a rounded copy, so the sum runs as one FMA chain from the mean.
template <typename Layout>
__device__ __forceinline__ void marker_af(const uint8_t *blob, const Layout &Y, const double *pc1, const double *pc2,
                                          int lane, double &af1, double &af2) {
  if (Y.panel_fp64) {
    const double *ud = reinterpret_cast<const double *>(blob + Y.off_ud);
    double a1 = 0., a2 = 0.;
    for (uint32_t k = 0; k < Y.n_pc; ++k) {
      const double u = ud[k * 32 + lane];
      a1 = __dadd_rn(a1, __dmul_rn(u, pc1[k]));
      a2 = __dadd_rn(a2, __dmul_rn(u, pc2[k]));
    }
    const double m = reinterpret_cast<const double *>(blob + Y.off_mu)[lane];
    af1 = (a1 + m) * 0.5;
    af2 = (a2 + m) * 0.5;
  } else {
    const float *ud = reinterpret_cast<const float *>(blob + Y.off_ud);
    double a1 = (double)reinterpret_cast<const float *>(blob + Y.off_mu)[lane], a2 = a1;
    if constexpr (Layout::kFixed) {
#pragma unroll
      for (uint32_t k = 0; k < Layout::n_pc; ++k) {
        const double u = (double)ud[k * 32 + lane];
        a1 = fma(u, pc1[k], a1);
        a2 = fma(u, pc2[k], a2);
      }
    } else {
      for (uint32_t k = 0; k < Y.n_pc; ++k) {
        const double u = (double)ud[k * 32 + lane];
        a1 = fma(u, pc1[k], a1);
        a2 = fma(u, pc2[k], a2);
      }
    }
    af1 = a1 * 0.5;
    af2 = a2 * 0.5;
  }
}

// header, allele frequencies, genotype priors, diagonal pairs of the slice whose chunk 0 sits at `buf`
struct SliceHeader {
  uint32_t wr, wa, n_valid, fr, fa, tails;
};
template <typename Layout, typename JobT>
__device__ __forceinline__ void slice_begin(const uint8_t *buf, const Layout &Y, const JobT &J, int lane,
                                            SliceHeader &H, double (&acc)[kNumPairs], double &ldiag) {
  const uint4 hdr = *reinterpret_cast<const uint4 *>(buf);
  H.wr = hdr.x; H.wa = hdr.y; H.n_valid = hdr.z & 0xFFu; H.tails = hdr.z >> 8; H.fr = hdr.w & 0xFFFFu; H.fa = hdr.w >> 16;
  double af1, af2;
  if (Y.known_af) {
    af1 = af2 = reinterpret_cast<const double *>(buf + Y.off_kaf)[lane];  // h:251-252
  } else {
    marker_af(buf, Y, J.pc1, J.pc2, lane, af1, af2);
  }
  double gf[3], gf2[3];
  initial_gf(af1, Y.min_af, Y.max_af, gf);   // contaminating sample
  initial_gf(af2, Y.min_af, Y.max_af, gf2);  // intended sample
  const double *dg = reinterpret_cast<const double *>(buf + Y.off_diag);
  ldiag = dg[lane] * (gf[0] * gf2[0]) + dg[32 + lane] * (gf[1] * gf2[1]) + dg[64 + lane] * (gf[2] * gf2[2]);
  // h:307-311 weights GF[g1]*GF2[g2]: start each running product at its weight, so the marginal is just
  // ldiag + sum(acc) at the end
#pragma unroll
  for (int p = 0; p < kNumPairs; ++p) acc[p] = gf[pair_g1(p)] * gf2[pair_g2(p)];
}

// the word rows [t_lo, t_hi) of the slice, stored from `col` on (this lane's column): ref rows first,
// then alt rows; rows [0,fr) and [wr, wr+fa) are filler-free in every lane.
template <bool MERGED, typename QT>
__device__ __forceinline__ void slice_rows(const uint32_t *col, uint32_t t_lo, uint32_t t_hi, const SliceHeader &H,
                                           const double *lin, const QT &Q, double (&acc)[kNumPairs]) {
  if (t_hi > H.wr + H.wa) t_hi = H.wr + H.wa;
  if (t_hi < t_lo) t_hi = t_lo;
  auto clampu = [&](uint32_t x) { return x < t_lo ? t_lo : (x > t_hi ? t_hi : x); };
  const uint32_t a0 = clampu(H.fr), a1 = clampu(H.wr), a2 = clampu(H.wr + H.fa);
  // (a uniform tail is only used when its row lies in this window together with the run's end)
  eat_runs<MERGED>(col, a0 - t_lo, a1 - a0, a1 == H.wr ? (H.tails & 0xFu) : 0u, a2 - a1, t_hi - a2,
                   t_hi == H.wr + H.wa ? ((H.tails >> 4) & 0xFu) : 0u, lin, Q, acc);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// Warp w issues on SM sub-partition w % 4 and owns bin blockIdx.x*4 + w%4; warp kk = w/4 serves the rounds
// kk, 2kc-1-kk, 2kc+kk, ... of that bin (a snake over groups of kc rounds: rounds are sorted heaviest first,
// so every warp gets about the same work).  Two launch geometries:
//   latency     one evaluation per launch: one CTA per SM with 4*kc warps (kc <= 4), all TMA copies of a
//               bin's first kc rounds in flight at once;
//   throughput  many evaluations per launch (gridDim.y = evaluation): kc = 1, four-warp CTAs, four of them
//               co-resident per SM, every warp walks all rounds of its bin behind a double-buffered TMA pipeline.
//   ARGS         sample, round table and job parameters are read from the kernel arguments
//                (constant bank, uniform addresses); otherwise from HBM through L1.
//   HOST_REDUCE  every CTA publishes {partial, seq} into the host-mapped mailbox and the host adds
//                them in a fixed order; otherwise the last CTA to finish adds the partials on the device.
//   NPC          2 or 4: FixedLayout<NPC> (fp32 panel, default clamps, no known AF); 0: RuntimeLayout.
//   CHUNKED      blobs may be deeper than one shared-memory stage and are streamed in chunks.
template <bool ARGS, bool HOST_REDUCE, int NPC, bool CHUNKED>
__global__ void __launch_bounds__(kMaxThreads, 1)
llk_kernel(const __grid_constant__ LaunchArgs A) {
  using Layout = typename std::conditional<NPC != 0, FixedLayout<NPC>, RuntimeLayout>::type;
  extern __shared__ __align__(128) uint8_t s_buf[];  // [warp][n_buf][buf_bytes], then (kc > 1) the marginals
  __shared__ __align__(16) JobParams s_job;
  __shared__ double s_red[4];
  __shared__ __align__(8) uint64_t s_bar[kMaxWarps][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t job = blockIdx.y;
  auto stamp = [&](int k) {  // diagnostics: SM clock of thread 0 at stage k (slot 8/9: global timer at entry/exit)
    if (A.trace && threadIdx.x == 0) {
      A.trace[blockIdx.x * kTraceSlots + k] = (unsigned long long)clock64();
      if (k == 0 || k == 6) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        A.trace[blockIdx.x * kTraceSlots + (k == 0 ? 8 : 9)] = g;
      }
    }
  };
  stamp(0);

  // ---- per-CTA set-up --------------------------------------------------------------------------
  // The code of a launch arrives cold from L2 (about five cycles per instruction), so the order is: arm the
  // warp's mbarriers and fire the TMA bulk copy of its first blob with as few instructions as possible, then
  // publish the tables, then look for the second blob -- all of it under the HBM latency of the first.
  const SampleDev &S = (ARGS || !A.samples) ? A.sample : A.samples[job];
  const vb2::Round *rounds_tab = ARGS ? A.rounds : S.rounds;
  const bool active_cta = blockIdx.x < S.grid_x;  // eval_many: a sample may need fewer CTAs than the grid has
  const Layout Y(S);
  const uint32_t bin = blockIdx.x * vb2::kBinsPerCta + (warp & 3);
  const uint32_t kc = A.kc, kk = (uint32_t)(warp >> 2);
  const uint32_t n_rounds = active_cta ? S.n_rounds : 0u;
  const uint32_t chunk_rows = S.chunk_rows;
  const uint32_t n_buf = A.n_buf, buf_bytes = S.buf_bytes;
  const uint8_t *blob_base = S.blob;
  uint8_t *mybuf = s_buf + (size_t)warp * n_buf * buf_bytes;

  // A stage = one TMA bulk copy = one blob (CHUNKED: one chunk of a blob).
  struct Stage {
    uint32_t off16, rows, c, n_ch, r;
    bool valid;
  };
  // chunk 0 = header + panel + diag + the first chunk_rows word rows; chunk c >= 1 = the next rows
  auto issue = [&](const Stage &s, uint32_t b) {
    uint32_t off = 0, bytes;
    if (!CHUNKED) {
      bytes = Y.off_words + s.rows * 128u;
    } else if (s.c == 0) {
      bytes = Y.off_words + (s.rows < chunk_rows ? s.rows : chunk_rows) * 128u;
    } else {
      off = Y.off_words + s.c * chunk_rows * 128u;
      const uint32_t n = s.rows - s.c * chunk_rows;
      bytes = (n < chunk_rows ? n : chunk_rows) * 128u;
    }
    mbar_arrive_expect_tx(&s_bar[warp][b], bytes);
    bulk_g2s(mybuf + (size_t)b * buf_bytes, blob_base + ((uint64_t)s.off16 << 4) + off, bytes, &s_bar[warp][b]);
  };
  // Warp kk of a bin serves the rounds kk, 2kc-1-kk, 2kc+kk, ...: its first blob is the one of round kk.
  Stage cur, nxt;
  cur.valid = nxt.valid = false;
  if (kk < n_rounds) {
    const vb2::Round R = rounds_tab[kk];
    if (bin - R.first_bin < R.count) {  // unsigned: also false when bin < first_bin
      cur.off16 = (uint32_t)((R.base + (uint64_t)(bin - R.first_bin) * R.stride) >> 4);
      cur.rows = R.rows;
      cur.c = 0;
      cur.n_ch = (!CHUNKED || R.rows <= chunk_rows) ? 1u : (R.rows + chunk_rows - 1) / chunk_rows;
      cur.r = kk;
      cur.valid = true;
    }
  }
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    mbar_fence_init();
    if (cur.valid) issue(cur, 0);
  }
  __syncwarp();
  stamp(7);

  // tables: Phred errors (kernel arguments -> shared memory), this evaluation's parameters, neutral marginals
  if (threadIdx.x < 256) s_e[threadIdx.x] = threadIdx.x < (uint32_t)kPhredArgs ? A.phred[threadIdx.x] : 1.0;
  if (blockDim.x < 256 && threadIdx.x < 128) s_e[threadIdx.x + 128] = 1.0;
  if (threadIdx.x < sizeof(JobParams) / sizeof(double))
    reinterpret_cast<double *>(&s_job)[threadIdx.x] =
        reinterpret_cast<const double *>(ARGS ? &A.jobs[job] : &A.jobs_dev[job])[threadIdx.x];
  // kc > 1: the marginals of a bin's slices meet in shared memory, [round][bin of the CTA][lane] (see below)
  double *s_L = reinterpret_cast<double *>(s_buf + (size_t)(blockDim.x >> 5) * n_buf * buf_bytes);
  if (kc > 1) {
#pragma unroll 1
    for (uint32_t i = threadIdx.x; i < n_rounds * 128u; i += blockDim.x) s_L[i] = 1.0;
  }

  // The rest of this warp's blobs: lane i of the item table holds round rbase + i -- whether this warp serves it
  // and this bin owns a blob in it, and where that blob is (all blobs of a round have one stride, so the address
  // is arithmetic on the round table; no per-blob descriptor is ever loaded).
  uint32_t rbase = 0, mask = 0, it_off16 = 0, it_rows = 0;
  auto load_table = [&]() {
    const uint32_t r = rbase + (uint32_t)lane;
    bool mine = false;
    if (r < n_rounds && r != kk) {
      const uint32_t m = r % (2u * kc);
      if (m == kk || m == 2u * kc - 1u - kk) {
        const vb2::Round R = rounds_tab[r];
        if (bin - R.first_bin < R.count) {
          mine = true;
          it_off16 = (uint32_t)((R.base + (uint64_t)(bin - R.first_bin) * R.stride) >> 4);
          it_rows = R.rows;
        }
      }
    }
    mask = __ballot_sync(0xFFFFFFFFu, mine);
  };
  uint32_t q_off16 = cur.off16, q_rows = cur.rows, q_c = cur.c, q_nch = cur.n_ch, q_r = cur.r;  // the issue cursor's blob
  auto advance = [&](Stage &s) {  // s = the next stage in this warp's order
    if (CHUNKED && q_c + 1 < q_nch) {
      ++q_c;
    } else {
      while (mask == 0) {
        rbase += 32;
        if (rbase >= n_rounds) {
          s.valid = false;
          return;
        }
        load_table();
      }
      const int i = __ffs((int)mask) - 1;
      mask &= mask - 1u;
      q_off16 = __shfl_sync(0xFFFFFFFFu, it_off16, i);
      q_rows = __shfl_sync(0xFFFFFFFFu, it_rows, i);
      q_r = rbase + (uint32_t)i;
      q_c = 0;
      q_nch = (!CHUNKED || q_rows <= chunk_rows) ? 1u : (q_rows + chunk_rows - 1) / chunk_rows;
    }
    s.off16 = q_off16; s.rows = q_rows; s.c = q_c; s.n_ch = q_nch; s.r = q_r; s.valid = true;
  };
  uint32_t ib = n_buf - 1, cb = 0, parity = 0;
  if (cur.valid && (CHUNKED || 2u * kc - 1u - kk < n_rounds)) {  // (otherwise round kk was this warp's only one)
    load_table();
    advance(nxt);
  }
  stamp(11);
  __syncthreads();
  stamp(1);

  // Sum over a BIN's markers of log(marginal), kept per lane as log(prod * 2^esum) + vsum: the marginals of the
  // bin's slices are multiplied up in round order (exponent split off after every factor) and ONE log per
  // lane is taken at the end.  kc == 1: the bin's only warp does that as it goes.  kc > 1: the bin's warps
  // leave their marginals in shared memory and warp (bin % 4) multiplies them up after the CTA barrier --
  // the same factors in the same order, so both launch geometries return the same bits.
  double vsum = 0.0, prod = 1.0;
  int esum = 0;
  auto combine = [&](double Lv) {  // Lv = marginal of one marker of this lane, or 1.0 (no marker / skipped)
    if (Lv > 1e-280) prod *= Lv;
    else vsum += cold_log(Lv);  // (sub)normal marginal of a very deep marker: no exponent tricks
    const int hi = __double2hiint(prod);  // prod in [1e-280, 2): positive and normal
    esum += (hi >> 20) - 1023;
    prod = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(prod));
  };
  if (cur.valid) {
    const double *lin = s_job.c0;  // c0[6] then c1[6] (contiguous in JobParams)
    Quad Q;
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) {
      const double c0 = s_job.c0[p], c1 = s_job.c1[p];
      Q.C0[p] = c0 * c0;
      Q.C1[p] = c0 * c1;
      Q.C2[p] = c1 * c1;
    }
    double acc[kNumPairs], ldiag = 0.;
    SliceHeader H{0, 0, 0, 0, 0, 0};
#pragma unroll 1
    while (cur.valid) {
      const Stage upcoming = nxt;  // fetched while `cur` is consumed (the host gives two buffers whenever a
      if (upcoming.valid) {        // warp can have more than one stage)
        __syncwarp();              // every lane finished reading the buffer about to be overwritten
        if (lane == 0) issue(upcoming, ib);
        advance(nxt);
        ib ^= 1u;
      }
      mbar_wait(&s_bar[warp][cb], (parity >> cb) & 1u);
      parity ^= 1u << cb;
      stamp(2);
      const uint8_t *buf = mybuf + (size_t)cb * buf_bytes;
      if (!CHUNKED || cur.c == 0) slice_begin(buf, Y, s_job, lane, H, acc, ldiag);
      bool last = true;
      if (!CHUNKED) {
        eat_runs<true>(reinterpret_cast<const uint32_t *>(buf + Y.off_words) + lane, H.fr, H.wr - H.fr, H.tails & 0xFu,
                       H.fa, H.wa - H.fa, (H.tails >> 4) & 0xFu, lin, Q, acc);
      } else {
        const uint32_t t_lo = cur.c * chunk_rows;
        slice_rows<true>(reinterpret_cast<const uint32_t *>(buf + (cur.c == 0 ? Y.off_words : 0u)) + lane, t_lo,
                         t_lo + chunk_rows, H, lin, Q, acc);
        last = cur.c + 1 == cur.n_ch;
      }
      if (last) {
        // h:307-311: markerLK = sum over the nine pairs (the running products carry their weights, the
        // diagonal pairs were folded into ldiag); markers with markerLK <= 0 (or NaN) are skipped (h:310)
        const double L = ldiag + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + (acc[4] + acc[5]));
        const double Lv = ((uint32_t)lane < H.n_valid && L > 0) ? L : 1.0;
        if (kc == 1) combine(Lv);
        else s_L[(cur.r * 4u + (uint32_t)(warp & 3)) * 32u + lane] = Lv;
      }
      cb ^= n_buf - 1u;
      cur = upcoming;
    }
  }
  stamp(3);
  if (A.trace && lane == 0)  // diagnostics: when the CTA's last warp left the read loop
    atomicMax(A.trace + blockIdx.x * kTraceSlots + 15, (unsigned long long)clock64());
  if (kc > 1) {
    __syncthreads();  // every marginal of the CTA's four bins is in shared memory
    stamp(4);
    if (warp >= 4) return;
    vsum = 0.0; prod = 1.0; esum = 0;
#pragma unroll 1
    for (uint32_t r = 0; r < n_rounds; ++r) combine(s_L[(r * 4u + (uint32_t)warp) * 32u + lane]);
  }
  stamp(12);
  vsum += fma((double)esum, kLn2, log(prod));
  // ---- fixed-order reduction: warp shuffle tree -> the CTA's four bins -> (host | last CTA) -----
#pragma unroll
  for (int o = 16; o; o >>= 1) vsum += __shfl_xor_sync(0xFFFFFFFFu, vsum, o);
  if (lane == 0) s_red[warp] = vsum;
  stamp(14);
  if (kc > 1) asm volatile("bar.sync 1, 128;" ::: "memory");  // (only warps 0..3 are still here)
  else __syncthreads();
  if (!active_cta || warp != 0) return;  // warp 0 finishes alone: no other warp waits for the grid-level work
  const double cta = ((s_red[0] + s_red[1]) + s_red[2]) + s_red[3];  // (every lane computes the same sum)
  stamp(5);
  if constexpr (HOST_REDUCE) {
    if (lane == 0) {
      Slot *slot = A.mbox + (size_t)job * S.grid_x + blockIdx.x;
      *reinterpret_cast<ulonglong2 *>(slot) = make_ulonglong2((unsigned long long)__double_as_longlong(cta), A.seq);
    }
    stamp(6);
  } else {
    const uint32_t pslot = A.slots ? A.slots[job] : job;
    const uint32_t grid_x = S.grid_x;
    double *part = S.partials + (size_t)pslot * grid_x;
    unsigned int ticket = 0;
    if (lane == 0) {
      part[blockIdx.x] = cta;
      __threadfence();
      ticket = atomicAdd(S.tickets + pslot, 1u);
    }
    ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
    if (ticket == grid_x - 1) {  // this CTA arrived last: add the partials in a fixed order
      __threadfence();
      double s = 0.0;
      for (uint32_t i = lane; i < grid_x; i += 32) s += __ldcg(part + i);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      if (lane == 0) {
        S.tickets[pslot] = 0u;  // ready for the next launch on this stream
        const double out = s + S.log_other_const;
        if (A.d_out) A.d_out[job] = out;
        if (A.mbox)
          *reinterpret_cast<ulonglong2 *>(A.mbox + job) =
              make_ulonglong2((unsigned long long)__double_as_longlong(out), A.seq);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the many-evaluations kernel (throughput geometry)
// ---------------------------------------------------------------------------------------------
// A launch of n evaluations is n * n_bins TASKS: task (j, b) = every round of bin b for evaluation j.  The grid
// is persistent -- four-warp CTAs, four co-resident per SM, 128 registers per thread -- and after one barrier
// that publishes the Phred table the warps never meet again: each warp pulls tasks from a queue in HBM (its
// first task is its own index, the rest come from an atomic counter whose next value is fetched one task
// ahead), walks the task's blobs behind a double-buffered TMA pipeline that runs across task boundaries, and
// leaves the task's sum in partials[j][b].  Which warp ran a task never shows in the result:
// llk_reduce_kernel, launched behind it, adds every evaluation's partials in the fixed order llk_kernel and
// the host use.  Everything the pipeline's issue side knows lives in shared memory (WarpCtl), so the read
// loops keep the registers.
constexpr int kRecRing = 4;  // task records resident per warp: the issue cursor's, <= 2 with unconsumed stages, 1 landing
struct WarpCtl {
  uint32_t next_task, n_rounds;
  uint32_t rbase;          // item table: entry i = round rbase + i
  uint32_t off16, rows, c, nch;  // CHUNKED: the blob being issued, chunk c of nch
  uint32_t done;           // queue exhausted
  uint32_t next_rec;       // ring slot of the next task's record (landing or landed)
  uint32_t rec_parity;     // bit s: mbarrier phase to wait for on ring slot s
  uint32_t it_off16[32], it_rows[32];
  uint32_t st_c[2], st_nch[2], st_chunk_rows[2];  // CHUNKED: which chunk of how many sits in each buffer
};
// What sits (or is landing) in one of a warp's two stage buffers -- a register, the same in every lane:
//   bit 31 = first stage of its task, bits 28..29 = ring slot of the task's record, bits 0..27 = the task's bin.
constexpr uint32_t kTagFirst = 0x80000000u;
__device__ __forceinline__ uint32_t tag_rec(uint32_t tag) { return (tag >> 28) & 3u; }
__device__ __forceinline__ uint32_t tag_bin(uint32_t tag) { return tag & 0x0FFFFFFFu; }

#ifndef VB2_STREAM_CTAS_PER_SM
#define VB2_STREAM_CTAS_PER_SM 4
#endif
#ifndef VB2_STREAM_PIPELINED
#define VB2_STREAM_PIPELINED 1   // 0: the plain consume loop for every shape (A/B builds)
#endif
template <int NPC, bool CHUNKED>
__global__ void __launch_bounds__(128, VB2_STREAM_CTAS_PER_SM)
llk_stream_kernel(const __grid_constant__ LaunchArgs A) {
  using Layout = typename std::conditional<NPC != 0, FixedLayout<NPC>, RuntimeLayout>::type;
  extern __shared__ __align__(128) uint8_t s_buf[];  // [warp][2][stage_bytes]
  __shared__ __align__(16) TaskRec s_rec[4][kRecRing];  // per warp: a ring of task records (TMA destinations)
  __shared__ __align__(8) uint64_t s_bar[4][2], s_rbar[4][kRecRing];
  __shared__ __align__(16) WarpCtl s_ctl[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_bins_max = A.n_bins_max, n_tasks = A.n_jobs * n_bins_max;
  WarpCtl &W = s_ctl[warp];
  // the record of task t -> ring slot `slot` (lane 0)
  auto prefetch = [&](uint32_t t, uint32_t slot) {
    mbar_arrive_expect_tx(&s_rbar[warp][slot], (uint32_t)sizeof(TaskRec));
    bulk_g2s(&s_rec[warp][slot], A.recs + t / n_bins_max, (uint32_t)sizeof(TaskRec), &s_rbar[warp][slot]);
  };
  // ---- task queue: the first task is the warp's own index, the rest come from an atomic counter whose next
  // value is fetched one task ahead; the RECORD of the next task is fetched as soon as its number is known
  uint32_t fetched = 0;  // (lane 0) the task after W.next_task; the atomic is in flight while a task runs
  auto fetch = [&]() {
    if (lane == 0) fetched = atomicAdd(A.queue, 1u) + gridDim.x * 4u;
  };
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    for (int i = 0; i < kRecRing; ++i) mbar_init(&s_rbar[warp][i], 1);
    mbar_fence_init();
    const uint32_t t0 = blockIdx.x * 4u + (uint32_t)warp;
    if (t0 < n_tasks) prefetch(t0, 0);
    W.next_task = t0;
    W.n_rounds = 0; W.rbase = 0; W.c = 0; W.nch = 0; W.done = 0;
    W.next_rec = 0; W.rec_parity = 0;
  }
  fetch();
  for (int i = threadIdx.x; i < 256; i += 128) s_e[i] = i < kPhredArgs ? A.phred[i] : 1.0;
#ifdef VB2_PHASE_CLOCK
  // diagnostics build (tools/gpu_phase.sh): cycles every warp spends in each phase of a slice, summed into A.trace
  __shared__ unsigned long long s_ph[4][kTraceSlots];
  if (lane < kTraceSlots) s_ph[warp][lane] = 0ull;
  unsigned long long ph_t = 0;
#define VB2_PH_START() do { if (lane == 0) ph_t = (unsigned long long)clock64(); } while (0)
#define VB2_PH(k) do { if (lane == 0) { const unsigned long long n_ = (unsigned long long)clock64(); s_ph[warp][k] += n_ - ph_t; ph_t = n_; } } while (0)
#define VB2_PH_COUNT(k) do { if (lane == 0) s_ph[warp][k] += 1ull; } while (0)
#else
#define VB2_PH_START() do { } while (0)
#define VB2_PH(k) do { } while (0)
#define VB2_PH_COUNT(k) do { } while (0)
#endif
  __syncthreads();  // the only CTA-wide barrier

  const uint32_t stage_bytes = A.stage_bytes;
  uint8_t *mybuf = s_buf + (size_t)warp * 2u * stage_bytes;

  // v = this lane's share of a task (record R, bin): xor-shuffle tree over the lanes -> partials[pslot][bin]
  auto store_partial = [&](const TaskRec &R, uint32_t bin, double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) R.S.partials[(size_t)R.pslot * (4u * R.S.grid_x) + bin] = v;
  };

  // ---- issue side ----------------------------------------------------------------------------------
  // Everything a stage needs lives in registers that hold the same value in every lane: the items of the current
  // table still to issue (mask), the issue cursor's task (its bin and record slot, whether it has issued anything
  // yet, the base of its sample's image) and the tags of the two stage buffers.  Shared memory (WarpCtl) holds the
  // item table and what is touched only when a task or a table runs out.
  uint32_t mask = 0, first = 0, i_bin = 0, i_rec = 0, tag0 = 0, tag1 = 0;
  const uint8_t *i_blob = nullptr;
  // mask == 0: the next 32 rounds of the task, or the next task; false = no more work
  auto next_table = [&]() -> bool {
    for (;;) {
      uint32_t rbase = W.rbase + 32u, n_rounds = W.n_rounds;
      if (rbase >= n_rounds) {  // the task's table is exhausted: take the next task
        if (first) store_partial(s_rec[warp][i_rec], i_bin, 0.0);  // (no blob at all: an empty bin still reports in)
        first = 0;
        const uint32_t t = W.next_task;
        if (t >= n_tasks) {
          __syncwarp();
          if (lane == 0) W.done = 1;
          __syncwarp();
          return false;
        }
        // the record of task t was requested when the previous task was taken (or at kernel entry)
        const uint32_t rec = W.next_rec, rpar = W.rec_parity;
        mbar_wait(&s_rbar[warp][rec], (rpar >> rec) & 1u);
        const uint32_t nt = __shfl_sync(0xFFFFFFFFu, fetched, 0);
        fetch();
        // the next record lands in a ring slot that is neither this task's nor that of a stage not yet consumed
        const uint32_t busy = (1u << rec) | (1u << tag_rec(tag0)) | (1u << tag_rec(tag1));
        const uint32_t nrec = (uint32_t)__ffs((int)(~busy & ((1u << kRecRing) - 1u))) - 1u;
        const TaskRec &R = s_rec[warp][rec];
        i_rec = rec;
        i_bin = t % n_bins_max;
        i_blob = R.S.blob;
        const bool active = i_bin < 4u * R.S.grid_x;  // eval_many: a sample may have fewer bins than the launch
        n_rounds = active ? R.S.n_rounds : 0u;
        __syncwarp();
        if (lane == 0) {
          if (nt < n_tasks) prefetch(nt, nrec);
          W.next_task = nt;
          W.n_rounds = n_rounds;
          W.rbase = 0u - 32u;
          W.next_rec = nrec;
          W.rec_parity = rpar ^ (1u << rec);
        }
        __syncwarp();
        if (!active) continue;
        first = 1;
        rbase = 0;
      }
      // item table of rounds [rbase, rbase + 32): lane i looks at round rbase + i
      const uint32_t r = rbase + (uint32_t)lane;
      bool mine = false;
      if (r < n_rounds) {
        const TaskRec &R = s_rec[warp][i_rec];
        const vb2::Round Rd = r < (uint32_t)kRecRounds ? R.rounds[r] : R.S.rounds[r];
        if (i_bin - Rd.first_bin < Rd.count) {  // unsigned: also false when bin < first_bin
          mine = true;
          W.it_off16[lane] = (uint32_t)((Rd.base + (uint64_t)(i_bin - Rd.first_bin) * Rd.stride) >> 4);
          W.it_rows[lane] = Rd.rows;
        }
      }
      __syncwarp();
      if (lane == 0) W.rbase = rbase;
      mask = __ballot_sync(0xFFFFFFFFu, mine);  // (also orders the table's stores before every lane's reads)
      if (mask) return true;
    }
  };
  // make sure there is an item to issue; false when the queue is exhausted (the rare path, kept out of line)
  auto prepare = [&]() -> bool {
    if (mask != 0) return true;
    return !W.done && next_table();
  };
  // Issue the next item (mask != 0) into buffer b; returns the buffer's tag.  Branch-free: every lane computes the
  // address, lane 0's copy is predicated.  The caller has made sure that every lane is done reading buffer b.
  auto issue_item = [&](uint32_t b) -> uint32_t {
    const int i = __ffs((int)mask) - 1;
    mask &= mask - 1u;
    const uint32_t off16 = W.it_off16[i], rows = W.it_rows[i];
    uint32_t off_words;
    if constexpr (Layout::kFixed) off_words = Layout::off_words;
    else off_words = s_rec[warp][i_rec].S.off_words;
    uint32_t bytes = off_words + rows * 128u;  // chunk 0 = header + panel + diag + the first chunk_rows word rows
    if (CHUNKED) {
      const uint32_t chunk_rows = s_rec[warp][i_rec].S.chunk_rows;
      const uint32_t nch = rows <= chunk_rows ? 1u : (rows + chunk_rows - 1) / chunk_rows;
      if (nch > 1) bytes = off_words + chunk_rows * 128u;
      __syncwarp();
      if (lane == 0) {
        W.off16 = off16; W.rows = rows; W.c = 0; W.nch = nch;
        W.st_c[b] = 0; W.st_nch[b] = nch; W.st_chunk_rows[b] = chunk_rows;
      }
      __syncwarp();
    }
    bulk_g2s_if(lane == 0, mybuf + (size_t)b * stage_bytes, i_blob + ((uint64_t)off16 << 4), bytes, &s_bar[warp][b]);
    const uint32_t tag = (first ? kTagFirst : 0u) | (i_rec << 28) | i_bin;
    first = 0;
    return tag;
  };
  // CHUNKED: the next chunk of the blob being issued, if any
  auto issue_chunk = [&](uint32_t b, uint32_t &tag) -> bool {
    if (!(W.c + 1 < W.nch)) return false;
    __syncwarp();
    if (lane == 0) {
      const SampleDev &S = s_rec[warp][i_rec].S;
      const uint32_t c = W.c + 1, chunk_rows = S.chunk_rows, rows = W.rows;
      const uint32_t off = S.off_words + c * chunk_rows * 128u, n = rows - c * chunk_rows;
      const uint32_t bytes = (n < chunk_rows ? n : chunk_rows) * 128u;
      mbar_arrive_expect_tx(&s_bar[warp][b], bytes);
      bulk_g2s(mybuf + (size_t)b * stage_bytes, i_blob + ((uint64_t)W.off16 << 4) + off, bytes, &s_bar[warp][b]);
      W.c = c;
      W.st_c[b] = c; W.st_nch[b] = W.nch; W.st_chunk_rows[b] = chunk_rows;
    }
    __syncwarp();
    tag = (i_rec << 28) | i_bin;
    return true;
  };
  // the next stage of this warp's task sequence into buffer b (all lanes are done with it); false: nothing left
  auto produce = [&](uint32_t b) -> bool {
    uint32_t tag = 0;
    if (!(CHUNKED && issue_chunk(b, tag))) {
      if (!prepare()) return false;
      __syncwarp();
      tag = issue_item(b);
    }
    if (b) tag1 = tag;
    else tag0 = tag;
    return true;
  };

  // ---- consume ------------------------------------------------------------------------------------
  uint32_t in_flight = 0, cb = 0, parity = 0;
  if (produce(0)) {
    in_flight = 1;
    if (produce(1)) in_flight = 2;
  }
  double vsum = 0.0, prod = 1.0;  // the running task: sum of log(marginal) = log(prod * 2^esum) + vsum
  int esum = 0;
  auto combine = [&](double Lv) {
    if (Lv > 1e-280) prod *= Lv;
    else vsum += cold_log(Lv);
    const int hi = __double2hiint(prod);
    esum += (hi >> 20) - 1023;
    prod = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(prod));
  };
  auto task_value = [&]() { return vsum + fma((double)esum, kLn2, log(prod)); };
  uint32_t c_bin = 0;
  const TaskRec *c_rec = &s_rec[warp][0];  // the record of the task being consumed
  StreamQuad Q;
  auto load_quad = [&]() {
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) {
      const double c0 = c_rec->J.c0[p], c1 = c_rec->J.c1[p];
      Q.C0[p] = c0 * c0;
      Q.C1[p] = c0 * c1;
      Q.C2[p] = c1 * c1;
    }
  };
  double acc[kNumPairs], ldiag = 0.;
  SliceHeader H{0, 0, 0, 0, 0, 0};
  VB2_PH_START();

  if constexpr (Layout::kFixed && !CHUNKED && VB2_STREAM_PIPELINED != 0) {
    // ---- the common shape, software-pipelined: while the warp still holds the products of slice n it (a) takes
    // the marginal of slice n, (b) sets up slice n+1 (header, allele frequencies, priors, starting weights) from
    // the other buffer and (c) issues slice n+2 into the buffer slice n has just left -- three independent
    // dependency chains in ONE basic block, so they overlap instead of queueing behind each other.
    const Layout Y(A.sample);
    if (in_flight) {
      mbar_wait(&s_bar[warp][0], 0u);
      parity = 1u;
      c_rec = &s_rec[warp][tag_rec(tag0)];
      c_bin = tag_bin(tag0);
      load_quad();
      slice_begin(mybuf, Y, c_rec->J, lane, H, acc, ldiag);
      VB2_PH_COUNT(9);
      for (;;) {
        VB2_PH(0);
        VB2_PH_COUNT(8);
        const uint8_t *buf = mybuf + (size_t)cb * stage_bytes;
        eat_runs<false>(reinterpret_cast<const uint32_t *>(buf + Y.off_words) + lane, H.fr, H.wr - H.fr, H.tails & 0xFu,
                        H.fa, H.wa - H.fa, (H.tails >> 4) & 0xFu, c_rec->J.c0, Q, acc);
        __syncwarp();  // every lane is done with buffer cb
        VB2_PH(4);
        const bool more = in_flight > 1;  // slice n+1 is in (or on its way into) the other buffer
        const bool can = prepare();       // there is a slice n+2 to issue
        const uint32_t ob = cb ^ 1u, tagN = cb ? tag0 : tag1;
        if (more && !mbar_test(&s_bar[warp][ob], (parity >> ob) & 1u)) mbar_wait(&s_bar[warp][ob], (parity >> ob) & 1u);
        VB2_PH(2);
        // h:307-311, as in llk_kernel
        const double L = ldiag + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + (acc[4] + acc[5]));
        const double Lv = ((uint32_t)lane < H.n_valid && L > 0) ? L : 1.0;
        const TaskRec *recN = (tagN & kTagFirst) ? &s_rec[warp][tag_rec(tagN)] : c_rec;
        if (more && can) {  // steady state
          slice_begin(mybuf + (size_t)ob * stage_bytes, Y, recN->J, lane, H, acc, ldiag);
          const uint32_t t = issue_item(cb);
          if (cb) tag1 = t;
          else tag0 = t;
          combine(Lv);
        } else {            // the warp's last slices
          combine(Lv);
          if (more) slice_begin(mybuf + (size_t)ob * stage_bytes, Y, recN->J, lane, H, acc, ldiag);
          if (can) {
            const uint32_t t = issue_item(cb);
            if (cb) tag1 = t;
            else tag0 = t;
          }
        }
        VB2_PH(3);
        in_flight += (can ? 1u : 0u) - 1u;
        if (!more) break;  // (then nothing was issued either: the queue ran dry before the pipeline did)
        parity ^= 1u << ob;
        if (tagN & kTagFirst) {  // slice n+1 opens a new task: close this one, switch to the new parameters
          VB2_PH_COUNT(9);
          store_partial(*c_rec, c_bin, task_value());
          c_rec = recN;
          c_bin = tag_bin(tagN);
          vsum = 0.0; prod = 1.0; esum = 0;
          load_quad();
          VB2_PH(1);
        }
        cb = ob;
      }
      store_partial(*c_rec, c_bin, task_value());
    }
  } else {
    Layout Y(A.sample);
    bool have_task = false;
    while (in_flight) {
      const uint32_t tag = cb ? tag1 : tag0;
      if (tag & kTagFirst) {  // a new task: close the previous one, switch to this evaluation's parameters (on chip)
        VB2_PH_COUNT(9);
        if (have_task) store_partial(*c_rec, c_bin, task_value());
        have_task = true;
        c_rec = &s_rec[warp][tag_rec(tag)];
        c_bin = tag_bin(tag);
        vsum = 0.0; prod = 1.0; esum = 0;
        load_quad();
        if (!Layout::kFixed) Y = Layout(c_rec->S);
        VB2_PH(1);
      }
      const JobParams &J = c_rec->J;
      const double *lin = J.c0;  // c0[6] then c1[6] (contiguous in JobParams)
      VB2_PH(0);
      mbar_wait(&s_bar[warp][cb], (parity >> cb) & 1u);
      parity ^= 1u << cb;
      VB2_PH(2);
      VB2_PH_COUNT(8);
      const uint8_t *buf = mybuf + (size_t)cb * stage_bytes;
      bool last = true;
      if (!CHUNKED) {
        slice_begin(buf, Y, J, lane, H, acc, ldiag);
        eat_runs<false>(reinterpret_cast<const uint32_t *>(buf + Y.off_words) + lane, H.fr, H.wr - H.fr, H.tails & 0xFu,
                        H.fa, H.wa - H.fa, (H.tails >> 4) & 0xFu, lin, Q, acc);
      } else {
        const uint32_t d_c = W.st_c[cb], d_chunk_rows = W.st_chunk_rows[cb];
        if (d_c == 0) slice_begin(buf, Y, J, lane, H, acc, ldiag);
        const uint32_t t_lo = d_c * d_chunk_rows;
        slice_rows<false>(reinterpret_cast<const uint32_t *>(buf + (d_c == 0 ? Y.off_words : 0u)) + lane, t_lo,
                          t_lo + d_chunk_rows, H, lin, Q, acc);
        last = d_c + 1 == W.st_nch[cb];
      }
      if (last) {  // h:307-311, as in llk_kernel
        const double L = ldiag + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + (acc[4] + acc[5]));
        combine(((uint32_t)lane < H.n_valid && L > 0) ? L : 1.0);
      }
      // the buffer just read is free: refill it
      VB2_PH(6);
      __syncwarp();  // every lane is done with buffer cb
      --in_flight;
      if (produce(cb)) ++in_flight;
      cb ^= 1u;
      VB2_PH(7);
    }
    if (have_task) store_partial(*c_rec, c_bin, task_value());
  }
  if (lane == 0 && fetched == 0xFFFFFFFFu) __threadfence();  // (retires the atomic still in flight)
#ifdef VB2_PHASE_CLOCK
  __syncwarp();
  if (A.trace && lane < kTraceSlots) atomicAdd(A.trace + lane, s_ph[warp][lane]);
#endif
}

// ---------------------------------------------------------------------------------------------
// the many-evaluations kernel, high-occupancy form (the common shapes: FixedLayout<2|4>, no chunking)
// ---------------------------------------------------------------------------------------------
// llk_stream_kernel keeps the read loop's 18 pair coefficients in per-thread registers (36 of its 128), and at 128
// registers only four warps fit per SM sub-partition: the FP64 pipe idles 40 % of the time because a warp spends
// half of a slice in the latency-bound phases around the read loops.  This kernel computes the same bits with
// 62 registers, so eight warps fit per sub-partition:
//   * everything an evaluation's arithmetic needs that is the same for all lanes -- the coefficients C0, C1, C2
//     and c0, c1, the PCs -- travels in the KERNEL ARGUMENTS (constant bank; CUDA 12.1+ allows 32,764 bytes) and is
//     read with a provably warp-uniform index, so the compiler fetches it into UNIFORM registers
//     (LDCU c[0][UR + off]) and feeds it to the DFMAs as uniform operands;
//   * for that the control flow around the read loops has to be provably uniform too: every value that steers it
//     and is the same in all lanes only by construction (the task number lane 0 drew from the queue, the job of
//     the stage being consumed, the slice's row counts read from the blob header) passes through a warp reduction
//     or a vote (REDUX / VOTEU write uniform registers), and the mbarrier wait leaves on a vote;
//   * no task records in shared memory: the few per-task facts (where the bin's blobs are, where its partial sum
//     goes) are read from the job's record in L2 when the task is taken.
// Tasks (job, bin) come from the same kind of queue as in llk_stream_kernel; every warp walks its task's blobs
// behind a two-stage TMA pipeline that runs across tasks, multiplies the marginals up per lane in round order and
// leaves the bin's sum in partials[job][bin]; llk_reduce_kernel adds them in the fixed order.  Same device
// functions, same order of operations: the same bits as llk_stream_kernel, llk_kernel and llk_session_kernel
// (tested).
struct FlowJob {  // one job's share of the kernel arguments
  double C1[kNumPairs], C2[kNumPairs];  // c0 c1 and c1^2: multiplicands of the read loop (one uniform operand per DFMA)
  double c0[kNumPairs], c1[kNumPairs];  // (contiguous: the ragged-row code indexes them as one array)
  double pc1[4], pc2[4];                // contaminant / intended PCs (NumPC <= 4)
};
// The read loop's view of a job: C0 = c0^2 is the ADDEND of the first DFMA of every pair, and a DFMA takes only one
// uniform operand, so C0 lives in (per-thread) registers -- formed by the same product as everywhere else -- while
// C1 and C2 stay in the constant bank / uniform registers.
#ifndef VB2_FLOW_PREFETCH_ROWS
#define VB2_FLOW_PREFETCH_ROWS 0   // 1: the read loop fetches row t+1 before the arithmetic of row t (A/B builds)
#endif
struct FlowQuad {
  static constexpr bool kPrefetchRows = VB2_FLOW_PREFETCH_ROWS != 0;
  static constexpr bool kDotAddress = true;
  double C0[kNumPairs];
  const double (&C1)[kNumPairs];
  const double (&C2)[kNumPairs];
  uint32_t sel[4];  // 8 << (8 * B): the byte dot product's selector of read B of a word (plain read loop)
};
#ifndef VB2_FLOW_JOBS
#define VB2_FLOW_JOBS 120
#endif
constexpr int kFlowJobs = VB2_FLOW_JOBS;
struct FlowArgs {
  const TaskRec *recs;     // job j of this launch = recs[j] (its sample, partial-sum slot, round table)
  uint32_t n_jobs;         // <= kFlowJobs
  uint32_t n_quads;        // bin quads (CTAs of a one-evaluation launch) per job in this launch
  uint32_t stage_bytes;    // bytes per shared-memory stage
  uint32_t pad_;
  unsigned int *queue;     // next task to hand out; zero between launches (llk_reduce_kernel rewinds it)
  double phred[kPhredArgs];
  FlowJob jobs[kFlowJobs];
};
static_assert(sizeof(FlowArgs) <= 32764, "kernel arguments must stay below 32,764 bytes");

// The register cap steers ptxas here: whether it keeps C1/C2 in uniform registers (wanted) or hoists them into vector
// registers for the short read loop and then spills the accumulators depends on its estimate of the pressure, and that
// changed with the code around the loop (an earlier version needed a 56-register cap; this one is in uniform mode at
// 64 with no spills, and the extra registers save re-derived addresses: 2.84 vs 2.89 us).  tools/gpu_ab.sh decides.
#ifndef VB2_FLOW_MIN_BLOCKS
#define VB2_FLOW_MIN_BLOCKS 8
#endif
constexpr uint32_t kFlowLast = 0x80000000u;  // stage tag: the last blob of its (job, bin)
template <int NPC>
__global__ void __launch_bounds__(128, VB2_FLOW_MIN_BLOCKS)
llk_flow_kernel(const __grid_constant__ FlowArgs F) {
  using Layout = FixedLayout<NPC>;
  extern __shared__ __align__(128) uint8_t s_buf[];  // [warp][2][stage_bytes]
  __shared__ __align__(8) uint64_t s_bar[4][2];
  __shared__ double *s_dst[4][2];  // where the partial sum of the task whose last blob sits in a stage goes
  __shared__ uint32_t s_sel[4];    // 8 << 8B: the byte dot product's selectors (see the plain read loop)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_jobs = F.n_jobs, n_quads = F.n_quads;
  // the next launch of the batch may move in as soon as this one leaves room (nothing here is an input of it)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 256; i += 128) s_e[i] = i < kPhredArgs ? F.phred[i] : 1.0;
  if (threadIdx.x < 4) s_sel[threadIdx.x] = 0x8u << (8u * threadIdx.x);
  __syncthreads();  // the only CTA-wide barrier

  const uint32_t stage_bytes = F.stage_bytes;
  uint8_t *mybuf = s_buf + (size_t)warp * 2u * stage_bytes;
#ifndef VB2_FLOW_OPAQUE_ADDR
#define VB2_FLOW_OPAQUE_ADDR 1   // 0: plain expressions, which ptxas re-derives from the thread index in every slice (A/B builds)
#endif
#if VB2_FLOW_OPAQUE_ADDR
  // shared-window addresses of this warp's stages and mbarriers, read back from shared memory so that ptxas keeps them
  // in two registers instead of re-deriving them from the thread index in every slice
  __shared__ uint32_t s_addr[4][2];
  if (lane == 0) {
    s_addr[warp][0] = smem_u32(mybuf);
    s_addr[warp][1] = smem_u32(&s_bar[warp][0]);
  }
  __syncwarp();
  const uint32_t buf0 = s_addr[warp][0], bar0 = s_addr[warp][1];
#else
  const uint32_t buf0 = smem_u32(mybuf), bar0 = smem_u32(&s_bar[warp][0]);  // (shared-window addresses, formed once)
#endif
  const Layout Y(F.recs[0].S);

  // ---- issue side: tasks (job, bin) from a counter in HBM; the blobs of a task in round order -------------------
  // A warp's first task is its own index, the following ones come from an atomic counter whose next value is
  // fetched one task ahead (the warps of a persistent grid do not progress evenly -- the schedulers favour some
  // -- so a static split leaves the pipe a third empty towards the end of a launch).  The task number reaches the
  // lanes through a warp reduction: a broadcast the compiler knows to be uniform.
  const uint32_t n_bins = n_quads * vb2::kBinsPerCta, n_tasks = n_jobs * n_bins;
  uint32_t fetched = blockIdx.x * 4u + (uint32_t)warp;  // (lane 0) the next task of this warp
  uint32_t i_job = 0;                // the job the cursor is in (uniform)
  uint32_t i_mask = 0;               // rounds of the cursor's task still to issue (bit r = round r; lane r holds it)
  const uint8_t *it_addr = nullptr;  // (lane r) where the bin's blob of round r is
  uint32_t it_rows = 0;              // (lane r) its word rows
  double *i_dst = nullptr;           // where the task's sum goes
  bool i_done = false;               // the counter ran past the last task
  // Lane r looks at round r of job `job` for bin `bin`; the ballot is the task's item mask.
  auto load_table = [&](uint32_t job, uint32_t bin) {
    const TaskRec &R = F.recs[job];
    const uint32_t grid_x = R.S.grid_x;
    const bool active = bin < vb2::kBinsPerCta * grid_x;  // eval_many: a sample may have fewer bins than the launch
    bool mine = false;
    if (active && (uint32_t)lane < R.S.n_rounds) {
      const vb2::Round Rd = lane < kRecRounds ? R.rounds[lane] : R.S.rounds[lane];
      if (bin - Rd.first_bin < Rd.count) {  // unsigned: also false when bin < first_bin
        mine = true;
        it_addr = R.S.blob + (Rd.base + (uint64_t)(bin - Rd.first_bin) * Rd.stride);
        it_rows = Rd.rows;
      }
    }
    i_dst = active ? R.S.partials + ((size_t)R.pslot * (vb2::kBinsPerCta * grid_x) + bin) : nullptr;
    i_mask = __ballot_sync(0xFFFFFFFFu, mine);
    if (i_mask == 0 && lane == 0 && i_dst) *i_dst = 0.0;  // (no blob at all: an empty bin still reports in)
  };
  // the next blob of this warp's sequence into stage b (every lane is done reading it); false: nothing left
  uint32_t tag0 = 0, tag1 = 0;  // what sits (or is landing) in the two stages: the job, kFlowLast
  auto produce = [&](uint32_t b) -> bool {
    while (i_mask == 0) {
      if (i_done) return false;
      const uint32_t t = __reduce_or_sync(0xFFFFFFFFu, lane == 0 ? fetched : 0u);
      if (t >= n_tasks) {
        i_done = true;
        return false;
      }
      if (lane == 0) fetched = atomicAdd(F.queue, 1u) + gridDim.x * 4u;  // (in flight while this task runs)
      i_job = t / n_bins;
      load_table(i_job, t - i_job * n_bins);
    }
    // the lane that holds the item fires its copy (no shuffles: every lane has its own blob's address and size)
    const int r = __ffs((int)i_mask) - 1;
    i_mask &= i_mask - 1u;
    const uint32_t tag = i_job | (i_mask == 0 ? kFlowLast : 0u);
    if (lane == r) {
      if (i_mask == 0) s_dst[warp][b] = i_dst;
      const uint32_t bytes = Layout::off_words + it_rows * 128u, bar = bar0 + 8u * b;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       buf0 + b * stage_bytes), "l"(it_addr), "r"(bytes), "r"(bar) : "memory");
    }
    if (b) tag1 = tag;
    else tag0 = tag;
    return true;
  };

  // ---- consume ---------------------------------------------------------------------------------------------------
  uint32_t in_flight = 0, cb = 0, parity = 0;
  if (produce(0)) {
    in_flight = 1;
    if (produce(1)) in_flight = 2;
  }
  double vsum = 0.0, prod = 1.0;  // the running task: sum of log(marginal) = log(prod * 2^esum) + vsum
  int esum = 0;
  // (read back from shared memory so that ptxas cannot fold them: they stay in four vector registers)
  const uint32_t sel0 = s_sel[0], sel1 = s_sel[1], sel2 = s_sel[2], sel3 = s_sel[3];
  while (__any_sync(0xFFFFFFFFu, in_flight != 0)) {
    // (the tag and the slice's row counts are the same in every lane; the reductions tell the compiler so)
    mbar_wait_warp(bar0 + 8u * cb, (parity >> cb) & 1u);
    parity ^= 1u << cb;
    const uint32_t tag = __reduce_or_sync(0xFFFFFFFFu, cb ? tag1 : tag0);
    const FlowJob &J = F.jobs[tag & 0xFFFFu];
#if VB2_FLOW_OPAQUE_ADDR
    const uint8_t *buf = static_cast<const uint8_t *>(__cvta_shared_to_generic(buf0 + cb * stage_bytes));
#else
    const uint8_t *buf = mybuf + (size_t)cb * stage_bytes;
#endif
    double acc[kNumPairs], ldiag;
    SliceHeader H;
    slice_begin(buf, Y, J, lane, H, acc, ldiag);
    const uint32_t u_rows = __reduce_or_sync(0xFFFFFFFFu, H.wr | (H.wa << 16));
    const uint32_t u_full = __reduce_or_sync(0xFFFFFFFFu, H.fr | (H.fa << 16));
    const uint32_t u_tails = __reduce_or_sync(0xFFFFFFFFu, H.tails);
    const uint32_t u_wr = u_rows & 0xFFFFu, u_wa = u_rows >> 16, u_fr = u_full & 0xFFFFu, u_fa = u_full >> 16;
    FlowQuad Q{{0., 0., 0., 0., 0., 0.}, J.C1, J.C2, {sel0, sel1, sel2, sel3}};
#pragma unroll
    for (int p = 0; p < kNumPairs; ++p) Q.C0[p] = J.c0[p] * J.c0[p];
    eat_runs<false>(reinterpret_cast<const uint32_t *>(buf + Y.off_words) + lane, u_fr, u_wr - u_fr, u_tails & 0xFu,
                    u_fa, u_wa - u_fa, (u_tails >> 4) & 0xFu, J.c0, Q, acc);
    // h:307-311, as in llk_kernel
    const double L = ldiag + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + (acc[4] + acc[5]));
    const double Lv = ((uint32_t)lane < H.n_valid && L > 0) ? L : 1.0;
    if (Lv > 1e-280) prod *= Lv;
    else vsum += cold_log(Lv);
    {
      const int hi = __double2hiint(prod);
      esum += (hi >> 20) - 1023;
      prod = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(prod));
    }
    if (tag & kFlowLast) {  // the task's last blob: its sum goes out
      double v = vsum + fma((double)esum, kLn2, log(prod));
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
      if (lane == 0) *s_dst[warp][cb] = v;  // (written by the lane that issued the blob, before the votes of the wait)
      vsum = 0.0; prod = 1.0; esum = 0;
    }
    __syncwarp();  // every lane is done with stage cb
    if (!produce(cb)) --in_flight;
    cb ^= 1u;
  }
}

// Behind llk_stream_kernel on the same stream: evaluation j's partials -> d_out[j] / mailbox slot j, in the fixed
// order of llk_kernel (the four bins of a CTA, those lane-strided over the CTAs, then a tree); rewinds the queue.
__global__ void __launch_bounds__(32, 1) llk_reduce_kernel(const __grid_constant__ LaunchArgs A) {
  const uint32_t job = blockIdx.x, lane = threadIdx.x;
  const SampleDev &S = A.recs[job].S;
  const uint32_t pslot = A.recs[job].pslot;
  const uint32_t grid_x = S.grid_x;
  const double *part = S.partials + (size_t)pslot * (4u * grid_x);
  double s = 0.0;
  for (uint32_t c = lane; c < grid_x; c += 32) {
    const double p0 = part[4u * c], p1 = part[4u * c + 1], p2 = part[4u * c + 2], p3 = part[4u * c + 3];
    s += ((p0 + p1) + p2) + p3;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) {
    const double out = s + S.log_other_const;
    if (A.d_out) A.d_out[job] = out;
    if (A.mbox)
      *reinterpret_cast<ulonglong2 *>(A.mbox + job) = make_ulonglong2((unsigned long long)__double_as_longlong(out), A.seq);
    // the task counters of the launches this batch was cut into (llk_stream_kernel: one; llk_flow_kernel: one per kFlowJobs jobs)
    if (job < (A.n_jobs + (uint32_t)kFlowJobs - 1u) / (uint32_t)kFlowJobs) A.queue[job] = 0u;
  }
  if (A.peer) {  // ---- fused with the collective: this shard's sum goes straight into every rank's buffer (NVLink stores)
    const PeerDev &P = *A.peer;
    const double out = __shfl_sync(0xFFFFFFFFu, s, 0) + S.log_other_const;
    const size_t bank = (size_t)(A.peer_seq & 1ull);
    if (lane < P.world)
      asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(P.vals[lane] + (bank * P.world + P.rank) * VB2_MAX_BATCH + job), "d"(out) : "memory");
    __threadfence_system();
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(P.ticket, 1u);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    if (t == gridDim.x - 1) {  // the launch's last job is out: stamp every rank's flag for this shard
      __threadfence_system();
      if (lane == 0) *P.ticket = 0u;
      if (lane < P.world)
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(P.flags[lane] + bank * P.world + P.rank), "l"(A.peer_seq) : "memory");
    }
  }
}

// Behind llk_reduce_kernel (peer mode): wait until every rank's shard sums of this launch have landed in THIS rank's
// buffer, then add them in rank order (the same order on every rank: the same bits everywhere).
__global__ void __launch_bounds__(256, 1) llk_gather_kernel(PeerDev P, unsigned long long seq, uint32_t n_jobs, double *d_out,
                                                           unsigned long long patience) {
  __shared__ int s_ok;
  const size_t bank = (size_t)(seq & 1ull);
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < P.world) {
    const unsigned long long *flag = P.flags[P.rank] + bank * P.world + threadIdx.x;
    const unsigned long long t0 = (unsigned long long)clock64();
    unsigned long long v;
    do {
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    } while (v != seq && (unsigned long long)clock64() - t0 < patience);
    if (v != seq) s_ok = 0;  // a rank never arrived: poison the results instead of spinning for ever
  }
  __syncthreads();
  const double *vals = P.vals[P.rank] + bank * P.world * VB2_MAX_BATCH;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_jobs; j += gridDim.x * blockDim.x) {
    double sum = 0.0;
    for (uint32_t r = 0; r < P.world; ++r) {
      double v;
      asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(vals + (size_t)r * VB2_MAX_BATCH + j) : "memory");
      sum += v;
    }
    d_out[j] = s_ok ? sum : __longlong_as_double(0x7FF8000000000000ll);
  }
}

// ---------------------------------------------------------------------------------------------
// the evaluation session: a resident kernel with the sample in shared memory
// ---------------------------------------------------------------------------------------------
// The simplex search evaluates ONE sample several hundred times, each evaluation waiting for the previous
// result, so a launch per evaluation pays the launch path of the part (~8 us from cudaLaunchKernel to a
// host-visible result for an empty kernel) every time.  A session launches llk_session_kernel ONCE: one CTA per
// SM, the geometry of llk_kernel, but every warp keeps ALL its blobs in shared memory (a 100k x 30x sample is
// ~50 KB per SM) and the CTA then serves evaluations until told to stop:
//   doorbell   host-mapped chunks {payload, seq}: the host writes the evaluation's coefficients and PCs and
//              stamps every 16-byte chunk with the sequence number; warp 0 of CTA 0 polls all chunks with one
//              coalesced load (one PCIe round trip) until every chunk carries the number it waits for and
//              forwards them to a copy in HBM, which the other CTAs poll in L2 (148 CTAs polling host memory
//              would saturate the PCIe read path: measured 276 us per evaluation);
//   compute    slice_begin / eat_runs / the marginals' meeting exactly as in llk_kernel (same bits), with no
//              HBM or L2 traffic at all;
//   answer     {CTA partial, seq} into the host mailbox; the host adds the partials in llk_kernel's order.
// The kernel leaves on the exit doorbell, or by itself after idle_cycles without a doorbell (so a host that
// went away cannot leave the GPU spinning); the host notices and relaunches.
constexpr int kMaxBellChunks = 12 + 2 * VB2_MAX_PC + 1;  // c0[6], c1[6], pc_contam[k], pc_intended[k], control
constexpr int kMaxSessionItems = 4;                  // blobs a warp may hold resident
constexpr unsigned long long kBellExit = ~0ull;
// A doorbell whose control chunk carries this flag (next to the sequence number in every other chunk) starts a
// simplex search on the device instead of one evaluation (vb2_llk_minimize).
constexpr unsigned long long kCtlMinimize = 1ull << 62;
constexpr unsigned long long kCtlDeviceMailbox = 1ull;  // relay only: this evaluation's partials go to the L2 mailbox

struct __align__(16) BellChunk {
  double payload;
  unsigned long long seq;
};

// ---- the simplex search on the device (vb2_llk_minimize) ---------------------------------------------------------
// Request and result live in host-mapped memory; the state lives in the shared memory of the first CTA.
constexpr int kMinDim = VB2_MIN_MAX_DIM;
constexpr int kMinPc = 4;
struct MinRequest {   // host -> device
  int32_t dim, n_pc, alpha_from, pad_;
  int32_t pc1_from[kMinPc], pc2_from[kMinPc];
  double pc1_fixed[kMinPc], pc2_fixed[kMinPc], alpha_fixed;
  double start[kMinDim], scale, ftol, llk1;
  long long cycle_max;
};
struct MinResult {    // device -> host; `done` is written last and carries the sequence number of the last evaluation
  double fmin, point[kMinDim], llk1, best_pc1[kMinPc], best_pc2[kMinPc], best_alpha;
  long long evals, cycle_count;
  int32_t converged, improved;
  unsigned long long done;
};
enum NmPhase : int { kNmIdle = 0, kNmInit, kNmReflect, kNmExpand, kNmContract, kNmShrink };
struct NmState {
  MinRequest rq;
  int phase, i, ilo, ihi, inhi, improved;
  int req_row;  // the point being evaluated: simplex row, or -1 = ptry
  int pad_;
  long long cycle_count, evals;
  double fmin, ysave, llk1, best_alpha;
  double simplex[kMinDim + 1][kMinDim], y[kMinDim + 1], psum[kMinDim], ptry[kMinDim];
  double best_pc1[kMinPc], best_pc2[kMinPc];
  double cur_pc1[kMinPc], cur_pc2[kMinPc], cur_alpha;  // ComputeMixLLKs' arguments of the evaluation in flight
  double fac[3], fac2[3];  // MathGenMin.cpp:427: (1 - factor) / dim and factor - that, for factor = -1, 2, 0.5
};

// AmoebaMinimizer (MathGenMin.cpp:326-443) as a resumable machine run by ONE WARP: nm_resume() takes the value of the
// point it asked for last and either asks for the next one (true: N.req_row) or finishes (false).  Lane j owns
// component j of every vector (the element-wise loops of statgen/MathVector.cpp:123-176 become one instruction per
// lane); the scalars live in shared memory, written by lane 0 and read by every lane into registers at entry, so the
// control flow is uniform and the decisions are taken on registers.  Every operation is the reference's, rounded
// where the reference rounds (no contraction: __dadd_rn / __dmul_rn / __ddiv_rn), so the simplex visits the
// reference's points when it is fed the reference's values.
__device__ __forceinline__ bool nm_resume(NmState &N, double f, int lane, int *converged) {
  const int dim = N.rq.dim, nvertex = dim + 1;
  const int phase = N.phase, i0 = N.i, ilo0 = N.ilo, ihi0 = N.ihi, inhi0 = N.inhi;
  const double ysave0 = N.ysave, fmin0 = N.fmin, ftol = N.rq.ftol, scale = N.rq.scale;
  long long cycles = N.cycle_count;
  const long long cycle_max = N.rq.cycle_max;
  double y[kMinDim + 1];  // every lane's copy of the vertex values
#pragma unroll
  for (int i = 0; i <= kMinDim; ++i) y[i] = N.y[i];
  // this lane's column of the simplex (lanes >= dim idle along)
  const bool mine = lane < dim;
  const int col = mine ? lane : 0;
  double psum = N.psum[col], ptry = N.ptry[col];
  const double start = N.rq.start[col];
  __syncwarp();  // every lane holds what it decides on; from here on the shared copies may be overwritten
  auto y_at = [&](int i) {  // (register array with a runtime index: a short select chain, no local memory)
    double v = y[0];
#pragma unroll
    for (int k = 1; k <= kMinDim; ++k) v = i == k ? y[k] : v;
    return v;
  };
  auto y_set = [&](int i, double v) {
#pragma unroll
    for (int k = 0; k <= kMinDim; ++k) y[k] = i == k ? v : y[k];
    if (lane == 0) N.y[i] = v;
  };
  auto compute_psum = [&]() {  // MathGenMin.cpp:347-349, :412-415
    double s = N.simplex[0][col];
    for (int m = 1; m <= dim; ++m) s = __dadd_rn(s, N.simplex[m][col]);
    psum = s;
  };
  // MathGenMin.cpp:425-433: the trial point through the face opposite vertex ihi
  auto make_try = [&](int ihi, int which) {  // which: 0 reflect (factor -1), 1 expand (2), 2 contract (0.5)
    const double fac = N.fac[which], fac2 = N.fac2[which];
    ptry = __dadd_rn(__dmul_rn(fac, psum), __dmul_rn(fac2, N.simplex[ihi][col]));
    if (mine) N.ptry[col] = ptry;
    if (lane == 0) N.req_row = -1;
  };
  // MathGenMin.cpp:434-442: the trial point replaces vertex ihi when it is better
  auto accept = [&](int ihi, double ytry) {
    if (ytry < y_at(ihi)) {
      y_set(ihi, ytry);
      psum = __dadd_rn(__dsub_rn(psum, N.simplex[ihi][col]), ptry);
      if (mine) N.simplex[ihi][col] = ptry;
    }
  };
  auto leave = [&](bool more) {  // write the lane-owned / lane-0 state back
    if (mine) N.psum[col] = psum;
    if (lane == 0) N.cycle_count = cycles;
    __syncwarp();
    return more;
  };
  auto ask_vertex = [&](int i, int next_phase) {
    if (lane == 0) { N.i = i; N.req_row = i; N.phase = next_phase; }
    return leave(true);
  };
  switch (phase) {
    case kNmIdle:  // start: the initial simplex, vertex by vertex (MathGenMin.cpp:335-345)
      if (lane < 3) {
        const double factor = lane == 0 ? -1.0 : (lane == 1 ? 2.0 : 0.5);
        const double fac = __ddiv_rn(__dsub_rn(1.0, factor), (double)dim);
        N.fac[lane] = fac;
        N.fac2[lane] = __dsub_rn(factor, fac);
      }
      if (mine) N.simplex[0][col] = __dadd_rn(start, col == 0 ? scale : 0.0);
      if (lane == 0) { N.fmin = 1.0e+100; N.ilo = N.ihi = N.inhi = 0; }
      return ask_vertex(0, kNmInit);
    case kNmInit: {
      y_set(i0, f);
      if (lane == 0 && f < fmin0) N.fmin = f;
      const int i = i0 + 1;
      if (i < nvertex) {
        if (mine) N.simplex[i][col] = i < dim ? __dadd_rn(start, col == i ? scale : 0.0) : start;
        return ask_vertex(i, kNmInit);
      }
      cycles = nvertex;
      compute_psum();
      break;
    }
    case kNmReflect: {
      const double y_ihi = y_at(ihi0);
      accept(ihi0, f);
      if (f <= y_at(ilo0)) {  // MathGenMin.cpp:392-394: expand
        make_try(ihi0, 1);
        if (lane == 0) N.phase = kNmExpand;
        return leave(true);
      }
      if (f >= y_at(inhi0)) {  // :395-399: contract (ysave = y[ihi] after the reflection was, or was not, accepted)
        make_try(ihi0, 2);
        if (lane == 0) { N.ysave = f < y_ihi ? f : y_ihi; N.phase = kNmContract; }
        return leave(true);
      }
      --cycles;  // :419
      break;
    }
    case kNmExpand:
      accept(ihi0, f);
      break;
    case kNmContract:
      accept(ihi0, f);
      if (!(f >= ysave0)) break;
      // :402-416: shrink everything towards the best vertex, re-evaluating vertex by vertex
      for (int i = 0; i < nvertex; ++i)
        if (i != ilo0) {
          if (mine) N.simplex[i][col] = __dmul_rn(__dadd_rn(N.simplex[i][col], N.simplex[ilo0][col]), 0.5);
          return ask_vertex(i, kNmShrink);
        }
      break;  // (dim == 0 cannot happen)
    case kNmShrink:
      y_set(i0, f);
      for (int i = i0 + 1; i < nvertex; ++i)
        if (i != ilo0) {
          if (mine) N.simplex[i][col] = __dmul_rn(__dadd_rn(N.simplex[i][col], N.simplex[ilo0][col]), 0.5);
          return ask_vertex(i, kNmShrink);
        }
      cycles += dim;
      compute_psum();
      break;
  }
  // the top of the loop (MathGenMin.cpp:357-390): order the vertices, test for convergence, reflect
  int ilo, ihi, inhi;
  if (y[0] > y[1]) { ilo = inhi = 1; ihi = 0; } else { ilo = inhi = 0; ihi = 1; }
  double v_lo = y_at(ilo), v_hi = y_at(ihi), v_nhi = y_at(inhi);
#pragma unroll
  for (int i = 2; i <= kMinDim; ++i)
    if (i < nvertex) {
      const double yi = y[i];
      if (yi <= v_lo) { ilo = i; v_lo = yi; }
      else if (yi > v_hi) { inhi = ihi; v_nhi = v_hi; ihi = i; v_hi = yi; }
      else if (yi > v_nhi) { inhi = i; v_nhi = yi; }
    }
  const double rtol = __ddiv_rn(2 * fabs(__dsub_rn(v_hi, v_lo)), __dadd_rn(__dadd_rn(fabs(v_hi), fabs(v_lo)), 3.0e-10));  // ZEPS
  if (lane == 0) { N.ilo = ilo; N.ihi = ihi; N.inhi = inhi; }
  if (rtol < ftol) {
    if (lane == 0) { N.fmin = v_lo; N.phase = kNmIdle; }
    *converged = 1;
    return leave(false);
  }
  if (cycles > cycle_max) {
    if (lane == 0) N.phase = kNmIdle;
    *converged = 0;
    return leave(false);
  }
  cycles += 2;
  make_try(ihi, 0);
  if (lane == 0) N.phase = kNmReflect;
  return leave(true);
}
// FullLLKFunc::Evaluate's unpacking of v (h:339-442) and fill_job's coefficients, rounded as the host rounds them:
// lane k takes PC k, lane p the genotype pair p
__device__ __forceinline__ void nm_job(NmState &N, JobParams &J, int lane) {
  const MinRequest &R = N.rq;
  const double *v = N.req_row >= 0 ? N.simplex[N.req_row] : N.ptry;
  if (lane < kMinPc) {
    const double a = lane < R.n_pc ? (R.pc1_from[lane] >= 0 ? v[R.pc1_from[lane]] : R.pc1_fixed[lane]) : 0.0;
    const double b = lane < R.n_pc ? (R.pc2_from[lane] >= 0 ? v[R.pc2_from[lane]] : R.pc2_fixed[lane]) : 0.0;
    N.cur_pc1[lane] = a; N.cur_pc2[lane] = b;
    J.pc1[lane] = a; J.pc2[lane] = b;
  }
  double alpha = R.alpha_fixed;
  if (R.alpha_from >= 0) {  // InvLogit, h:119-122
    const double e = exp(v[R.alpha_from]);
    alpha = __ddiv_rn(e, __dadd_rn(1., e));
  }
  if (lane == 0) N.cur_alpha = alpha;
  if (lane < kNumPairs) {
    const int g1 = lane < 2 ? 0 : (lane < 4 ? 1 : 2);
    const int g2 = lane == 0 ? 1 : lane == 1 ? 2 : lane == 2 ? 0 : lane == 3 ? 2 : lane == 4 ? 0 : 1;
    const double Eg1 = g1 == 0 ? 0.0 : (g1 == 1 ? 1.0 / 6.0 : 1.0 / 3.0), Eg2 = g2 == 0 ? 0.0 : (g2 == 1 ? 1.0 / 6.0 : 1.0 / 3.0);
    const double Ng1 = g1 == 0 ? 1.0 : (g1 == 1 ? 0.5 : 0.0), Ng2 = g2 == 0 ? 1.0 : (g2 == 1 ? 0.5 : 0.0);
    const double oma = __dsub_rn(1.0, alpha);
    const double e_mix = __dadd_rn(__dmul_rn(alpha, Eg1), __dmul_rn(oma, Eg2));
    const double n_mix = __dadd_rn(__dmul_rn(alpha, Ng1), __dmul_rn(oma, Ng2));
    J.c0[lane] = n_mix;
    J.c1[lane] = __dsub_rn(e_mix, n_mix);
  }
  __syncwarp();
}
// Evaluate's best-so-far bookkeeping (h:344-440): the free components of the best point over all evaluations
__device__ __forceinline__ void nm_track_best(NmState &N, double f, int lane) {
  const bool better = f < N.llk1;
  __syncwarp();
  if (better) {
    const MinRequest &R = N.rq;
    if (lane < kMinPc) {
      if (R.pc1_from[lane] >= 0) N.best_pc1[lane] = N.cur_pc1[lane];
      if (R.pc2_from[lane] >= 0) N.best_pc2[lane] = N.cur_pc2[lane];
    }
    if (lane == 0) {
      N.llk1 = f;
      N.improved = 1;
      if (R.alpha_from >= 0) N.best_alpha = N.cur_alpha;
    }
  }
  __syncwarp();
}

struct SessionArgs {
  SampleDev sample;
  const BellChunk *bell;  // device view of the host-mapped doorbell (polled by CTA 0 only: PCIe reads)
  BellChunk *relay;       // the same chunks in HBM: CTA 0 forwards the doorbell, the other CTAs poll this copy in L2
  Slot *mbox;             // device view of the host mailbox: slot [cta]
  Slot *dmbox;            // the inboxes in HBM/L2: slot [seq & 1][destination CTA][source CTA] (a search on the device)
  const MinRequest *min_req;  // device views of the host-mapped request / result of vb2_llk_minimize
  MinResult *min_res;
  MinRequest *min_req_dev;    // the request, forwarded into HBM by CTA 0 for the other CTAs
  unsigned long long first_seq, idle_cycles;
  uint32_t kc, n_items, n_chunks;  // n_chunks = 12 + 2 n_pc parameter chunks; the control chunk follows them
  unsigned long long *trace;  // diagnostics: [grid_x][kTraceSlots] SM clock of thread 0 at the stages of the LAST evaluation
  uint32_t null_eval;  // diagnostics (VB2_LLK_SESSION_NULL): answer without reading a single read -> the doorbell + mailbox round trip
  double phred[kPhredArgs];
  vb2::Round rounds[kMaxArgRounds];
};

// A search on the device (vb2_llk_minimize) runs in EVERY CTA at once: the simplex is a deterministic function of
// the values it is fed, so instead of one CTA proposing points and broadcasting them, every CTA's first warp keeps
// its own copy of the simplex, all-gathers the per-CTA partial sums of an evaluation through a two-bank mailbox in
// L2, adds them in the one fixed order and steps its copy -- the same bits everywhere, and one L2 hop per evaluation.
template <int NPC>
__global__ void __launch_bounds__(kSessionMaxWarps * 32, 1)
llk_session_kernel(const __grid_constant__ SessionArgs A) {
  using Layout = typename std::conditional<NPC != 0, FixedLayout<NPC>, RuntimeLayout>::type;
  extern __shared__ __align__(128) uint8_t s_buf[];  // [warp][n_items][buf_bytes], then the marginals
  __shared__ __align__(16) JobParams s_job;
  __shared__ double s_red[4];
  __shared__ __align__(8) uint64_t s_bar[kSessionMaxWarps];
  __shared__ uint32_t s_item_r[kSessionMaxWarps][kMaxSessionItems];
  __shared__ uint32_t s_stop, s_mode;
  __shared__ __align__(16) NmState s_nm;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SampleDev &S = A.sample;
  const Layout Y(S);
  const uint32_t bin = blockIdx.x * vb2::kBinsPerCta + (warp & 3);
  const uint32_t kc = A.kc, kk = (uint32_t)(warp >> 2);
  const uint32_t n_rounds = S.n_rounds, buf_bytes = S.buf_bytes;
  uint8_t *mybuf = s_buf + (size_t)warp * A.n_items * buf_bytes;
  double *s_L = reinterpret_cast<double *>(s_buf + (size_t)(blockDim.x >> 5) * A.n_items * buf_bytes);

  // ---- once: fetch this warp's blobs (rounds kk, 2kc-1-kk, 2kc+kk, ... in which the bin owns one) -----
  if (lane == 0) {
    mbar_init(&s_bar[warp], 1);
    mbar_fence_init();
    uint32_t n = 0, bytes_total = 0;
    uint32_t r = kk, odd = 0;
    while (r < n_rounds && n < (uint32_t)kMaxSessionItems) {
      const vb2::Round R = A.rounds[r];
      if (bin - R.first_bin < R.count) {  // unsigned: also false when bin < first_bin
        bytes_total += Y.off_words + R.rows * 128u;
        s_item_r[warp][n++] = r;
      }
      r += odd ? 2u * kk + 1u : 2u * kc - 1u - 2u * kk;
      odd ^= 1u;
    }
    for (uint32_t j = n; j < (uint32_t)kMaxSessionItems; ++j) s_item_r[warp][j] = 0xFFFFFFFFu;
    if (n) {
      mbar_arrive_expect_tx(&s_bar[warp], bytes_total);
      for (uint32_t j = 0; j < n; ++j) {
        const vb2::Round R = A.rounds[s_item_r[warp][j]];
        bulk_g2s(mybuf + (size_t)j * buf_bytes, S.blob + R.base + (uint64_t)(bin - R.first_bin) * R.stride,
                 Y.off_words + R.rows * 128u, &s_bar[warp]);
      }
    }
  }
  if (threadIdx.x < 256) s_e[threadIdx.x] = threadIdx.x < (uint32_t)kPhredArgs ? A.phred[threadIdx.x] : 1.0;
  if (blockDim.x < 256 && threadIdx.x < 128) s_e[threadIdx.x + 128] = 1.0;
  if (threadIdx.x == 0) {
    s_stop = 0u;
    s_mode = 0u;
    s_nm.phase = kNmIdle;
    s_nm.i = s_nm.ilo = s_nm.ihi = s_nm.inhi = 0;
    s_nm.rq.dim = 1;
  }
#pragma unroll 1
  for (uint32_t i = threadIdx.x; i < n_rounds * 128u; i += blockDim.x) s_L[i] = 1.0;  // neutral marginals
  __syncwarp();
  uint32_t n_items = 0;
  while (n_items < (uint32_t)kMaxSessionItems && s_item_r[warp][n_items] != 0xFFFFFFFFu) ++n_items;
  if (n_items) mbar_wait(&s_bar[warp], 0u);  // resident from here on

  auto stamp = [&](int k) {
    if (A.trace && threadIdx.x == 0) A.trace[blockIdx.x * kTraceSlots + k] = (unsigned long long)clock64();
  };
  // chunk i of the doorbell <-> JobParams: c0[6], c1[6], pc1[k] (at 12), pc2[k] (at 12 + VB2_MAX_PC)
  const uint32_t n_chunks = A.n_chunks, n_pc_chunks = S.n_pc;
  auto chunk_slot = [&](uint32_t i) { return i < 12u ? i : (i < 12u + n_pc_chunks ? i : i - n_pc_chunks + (uint32_t)VB2_MAX_PC); };
  const bool head = blockIdx.x == 0;  // the one CTA that talks to the host
  unsigned long long expected = A.first_seq;
  bool searching = false;  // (warp 0) a simplex search is running: the next point comes from this CTA's s_nm
#pragma unroll 1
  for (;;) {
    stamp(0);
    // ---- the next evaluation's parameters ----------------------------------------------------------------
    if (warp == 0) {
      uint32_t stop = 0, mode = 0;
      if (searching) {
        // the search proposes the point itself: Evaluate's unpacking + fill_job (no doorbell, no relay)
        nm_job(s_nm, s_job, lane);
        mode = (uint32_t)kCtlDeviceMailbox;
      } else {
        const unsigned long long t_idle = (unsigned long long)clock64();
        // Only the head CTA decides that the session has been idle for too long (and says so through the relay);
        // the others would only give up, much later, on a head that has gone away.
        const unsigned long long patience = head ? A.idle_cycles : A.idle_cycles * 64ull;
        const BellChunk *src = head ? A.bell : A.relay;
        unsigned long long ctl = 0ull;
        for (;;) {
          bool ok = true, bye = false;
          unsigned long long xsum = 0ull;
          ctl = 0ull;
#pragma unroll
          for (uint32_t base = 0; base < (uint32_t)kMaxBellChunks; base += 32) {
            const uint32_t i = base + (uint32_t)lane;
            if (i <= n_chunks) {
              // ONE 16-byte load per chunk: payload and stamp arrive together (a naturally aligned 16-byte access is
              // a single transaction on this part; the host's control chunk also carries a checksum of the
              // payloads, so a torn read could not pass for a doorbell)
              unsigned long long pay, seq;
              asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(pay), "=l"(seq) : "l"(src + i) : "memory");
              bye = bye || seq == kBellExit;
              if (i < n_chunks) {
                ok = ok && seq == expected;
                xsum ^= pay;
                if (seq == expected) reinterpret_cast<double *>(&s_job)[chunk_slot(i)] = __longlong_as_double((long long)pay);
              } else {  // the control chunk
                ok = ok && (seq & ~kCtlMinimize) == expected;
                ctl = seq & kCtlMinimize;
                if (head) xsum ^= pay;
              }
            }
          }
          if (__any_sync(0xFFFFFFFFu, bye)) { stop = 1; break; }
          if (__all_sync(0xFFFFFFFFu, ok)) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
              xsum ^= __shfl_xor_sync(0xFFFFFFFFu, xsum, o);
              ctl |= __shfl_xor_sync(0xFFFFFFFFu, ctl, o);
            }
            if (!head || (xsum & ~0xFFull) == 0ull) break;  // (host doorbell: the payloads' checksum, bits 8..63, must match)
          }
          if ((unsigned long long)clock64() - t_idle > patience) { stop = 1; break; }
        }
        if (!stop && (ctl & kCtlMinimize)) {
          // start a search: the request travels host -> head CTA -> HBM -> every other CTA
          constexpr uint32_t kWords = (uint32_t)(sizeof(MinRequest) / 8u);
          unsigned long long *dst = reinterpret_cast<unsigned long long *>(&s_nm.rq);
          if (head) {
            __threadfence_system();
            const volatile unsigned long long *rq = reinterpret_cast<const volatile unsigned long long *>(A.min_req);
            unsigned long long *fwd = reinterpret_cast<unsigned long long *>(A.min_req_dev);
            for (uint32_t w = lane; w < kWords; w += 32) {
              const unsigned long long x = rq[w];
              dst[w] = x;
              asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(fwd + w), "l"(x) : "memory");
            }
            __threadfence();
          } else {
            __threadfence();
            const unsigned long long *fwd = reinterpret_cast<const unsigned long long *>(A.min_req_dev);
            for (uint32_t w = lane; w < kWords; w += 32) {
              unsigned long long x;
              asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(x) : "l"(fwd + w) : "memory");
              dst[w] = x;
            }
          }
          if (lane == 0) {
            s_nm.phase = kNmIdle;
            s_nm.evals = 0;
            s_nm.improved = 0;
          }
          __syncwarp();
          if (lane == 0) s_nm.llk1 = s_nm.rq.llk1;
          int conv = 0;
          nm_resume(s_nm, 0.0, lane, &conv);  // (asks for the first vertex)
          searching = true;
        }
        if (head) {  // forward: the evaluation's chunks (or the start of a search), or the order to leave
          __syncwarp();
#pragma unroll
          for (uint32_t base = 0; base < (uint32_t)kMaxBellChunks; base += 32) {
            const uint32_t i = base + (uint32_t)lane;
            if (i <= n_chunks) {
              const unsigned long long pay = i < n_chunks
                  ? (unsigned long long)__double_as_longlong(reinterpret_cast<const double *>(&s_job)[chunk_slot(i)]) : 0ull;
              const unsigned long long seq = stop ? kBellExit : (i < n_chunks ? expected : (expected | (ctl & kCtlMinimize)));
              asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(A.relay + i), "l"(pay), "l"(seq) : "memory");
            }
          }
        }
        if (searching) {
          nm_job(s_nm, s_job, lane);
          mode = (uint32_t)kCtlDeviceMailbox;
        }
      }
      if (lane == 0) {
        if (stop) s_stop = 1u;
        s_mode = mode;
      }
      stamp(1);
    }
    __syncthreads();  // parameters (or the stop flag) published
    if (s_stop) return;
    const uint32_t mode = s_mode;
    stamp(2);

    // ---- the evaluation: every resident slice of this warp -------------------------------------------
    if (n_items && !A.null_eval) {
      const double *lin = s_job.c0;
      Quad Q;
#pragma unroll
      for (int p = 0; p < kNumPairs; ++p) {
        const double c0 = s_job.c0[p], c1 = s_job.c1[p];
        Q.C0[p] = c0 * c0;
        Q.C1[p] = c0 * c1;
        Q.C2[p] = c1 * c1;
      }
#pragma unroll 1
      for (uint32_t j = 0; j < n_items; ++j) {
        const uint8_t *buf = mybuf + (size_t)j * buf_bytes;
        double acc[kNumPairs], ldiag = 0.;
        SliceHeader H{0, 0, 0, 0, 0, 0};
        slice_begin(buf, Y, s_job, lane, H, acc, ldiag);
        eat_runs<true>(reinterpret_cast<const uint32_t *>(buf + Y.off_words) + lane, H.fr, H.wr - H.fr, H.tails & 0xFu,
                       H.fa, H.wa - H.fa, (H.tails >> 4) & 0xFu, lin, Q, acc);
        const double L = ldiag + ((acc[0] + acc[1]) + (acc[2] + acc[3]) + (acc[4] + acc[5]));  // h:307-311
        s_L[(s_item_r[warp][j] * 4u + (uint32_t)(warp & 3)) * 32u + lane] = ((uint32_t)lane < H.n_valid && L > 0) ? L : 1.0;
      }
    }
    stamp(3);
    if (A.trace && lane == 0) atomicMax(A.trace + blockIdx.x * kTraceSlots + 15, (unsigned long long)clock64());
    __syncthreads();  // every marginal of the CTA's four bins is in shared memory
    if (warp < 4) {
      double vsum = 0.0, prod = 1.0;
      int esum = 0;
#pragma unroll 1
      for (uint32_t r = 0; r < n_rounds; ++r) {  // the same factors in the same order as llk_kernel
        const double Lv = s_L[(r * 4u + (uint32_t)warp) * 32u + lane];
        if (Lv > 1e-280) prod *= Lv;
        else vsum += cold_log(Lv);
        const int hi = __double2hiint(prod);
        esum += (hi >> 20) - 1023;
        prod = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(prod));
      }
      vsum += fma((double)esum, kLn2, log(prod));
#pragma unroll
      for (int o = 16; o; o >>= 1) vsum += __shfl_xor_sync(0xFFFFFFFFu, vsum, o);
      if (lane == 0) s_red[warp] = vsum;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 0) {
        const uint32_t gx = S.grid_x;
        double cta = 0.0;
        if (lane == 0) cta = ((s_red[0] + s_red[1]) + s_red[2]) + s_red[3];
        if (!(mode & (uint32_t)kCtlDeviceMailbox)) {
          if (lane == 0) {
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(A.mbox + blockIdx.x),
                         "l"((unsigned long long)__double_as_longlong(cta)), "l"(expected) : "memory");
            stamp(6);
          }
        } else {
          // ---- all-gather, push style: this CTA's partial goes into EVERY CTA's inbox (148 posted 16-byte stores,
          // spread over the L2 slices), and every CTA then polls only its own inbox -- all CTAs polling one shared
          // mailbox would queue thousands of loads on the one L2 slice that holds it.  Two banks (by the parity of
          // the sequence number): a fast CTA's next partial must not overwrite the one a slow CTA still waits for.
          cta = __shfl_sync(0xFFFFFFFFu, cta, 0);
          Slot *bank = A.dmbox + (size_t)(expected & 1ull) * gx * gx;
          for (uint32_t dest = lane; dest < gx; dest += 32)
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(bank + (size_t)dest * gx + blockIdx.x),
                         "l"((unsigned long long)__double_as_longlong(cta)), "l"(expected) : "memory");
          stamp(6);
          // the partials are added in the order the host (and llk_kernel's last CTA) use -- lane-strided sums, then a
          // butterfly -- so every CTA steps its simplex on the same bits
          const Slot *inbox = bank + (size_t)blockIdx.x * gx;
          double sum = 0.0;
          for (uint32_t c0 = 0; c0 < gx; c0 += 256) {  // up to eight slots per lane in flight at once
            unsigned long long val[8], seq[8];
            for (;;) {
              bool all = true;
#pragma unroll
              for (uint32_t u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u * 32u + (uint32_t)lane;
                val[u] = 0ull; seq[u] = expected;
                if (c < gx)
                  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(val[u]), "=l"(seq[u]) : "l"(inbox + c) : "memory");
              }
#pragma unroll
              for (uint32_t u = 0; u < 8; ++u) all = all && seq[u] == expected;
              if (all) break;
            }
#pragma unroll
            for (uint32_t u = 0; u < 8; ++u)   // (c ascending per lane: the host's order)
              if (c0 + u * 32u + (uint32_t)lane < gx) sum += __longlong_as_double((long long)val[u]);
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
          const double f = 0 - (sum + S.log_other_const);  // h:344
          stamp(4);
          if (lane == 0) ++s_nm.evals;
          nm_track_best(s_nm, f, lane);
          int conv = 0;
          const bool more = nm_resume(s_nm, f, lane, &conv);
          stamp(5);
          if (!more) {  // finished everywhere; the head CTA reports: the result, then its stamp
            searching = false;
            if (head && lane == 0) {
              volatile MinResult *res = A.min_res;
              res->fmin = s_nm.fmin;
              // (without convergence AmoebaMinimizer::point is never assigned: it still is the starting point)
              for (int j = 0; j < kMinDim; ++j)
                res->point[j] = j < s_nm.rq.dim ? (conv ? s_nm.simplex[s_nm.ilo][j] : s_nm.rq.start[j]) : 0.0;
              res->llk1 = s_nm.llk1;
              for (int k = 0; k < kMinPc; ++k) { res->best_pc1[k] = s_nm.best_pc1[k]; res->best_pc2[k] = s_nm.best_pc2[k]; }
              res->best_alpha = s_nm.best_alpha;
              res->evals = s_nm.evals;
              res->cycle_count = s_nm.cycle_count;
              res->converged = conv;
              res->improved = s_nm.improved;
              __threadfence_system();
              res->done = expected;
            }
          }
        }
      }
    }
    ++expected;
  }
}

// Pick the instantiation for a launch.  spec = 2 / 4 when every sample of the launch has the FixedLayout<spec>
// shape, else 0; chunked = some blob is deeper than its shared-memory stage.
template <bool ARGS, bool HOST_REDUCE>
void launch_llk(dim3 grid, dim3 block, uint32_t smem, cudaStream_t stream, const LaunchArgs &A, int spec, bool chunked) {
  if (chunked) llk_kernel<ARGS, HOST_REDUCE, 0, true><<<grid, block, smem, stream>>>(A);
  else if (spec == 2) llk_kernel<ARGS, HOST_REDUCE, 2, false><<<grid, block, smem, stream>>>(A);
  else if (spec == 4) llk_kernel<ARGS, HOST_REDUCE, 4, false><<<grid, block, smem, stream>>>(A);
  else llk_kernel<ARGS, HOST_REDUCE, 0, false><<<grid, block, smem, stream>>>(A);
}
// CTAs of an llk_stream_kernel launch: as many as are co-resident (occupancy x SMs), never more than the tasks need.
template <int NPC, bool CHUNKED>
void launch_stream_as(uint32_t smem, cudaStream_t stream, const LaunchArgs &A, int sm_count) {
  static thread_local uint32_t cached_smem = ~0u, cached_per_sm = 0;  // (per instantiation and host thread)
  if (cached_smem != smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, llk_stream_kernel<NPC, CHUNKED>, 128, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    cached_per_sm = (uint32_t)per_sm;
    cached_smem = smem;
  }
  const uint32_t n_tasks = A.n_jobs * A.n_bins_max;
  const uint32_t grid = std::max(1u, std::min(cached_per_sm * (uint32_t)std::max(1, sm_count), (n_tasks + 3u) / 4u));
  llk_stream_kernel<NPC, CHUNKED><<<dim3(grid, 1, 1), dim3(128, 1, 1), smem, stream>>>(A);
}
bool flow_overlap() {  // VB2_FLOW_PDL=0: the launches of a batch strictly one after the other (A/B runs)
  const char *e = getenv("VB2_FLOW_PDL");
  return !(e && e[0] == '0');
}
// llk_flow_kernel over the jobs A.recs[0 .. A.n_jobs) (h_recs = the host copy of the same records): launches of at
// most kFlowJobs jobs, each with its jobs' parameters in the kernel arguments.
template <int NPC>
void launch_flow_as(uint32_t smem, cudaStream_t stream, const LaunchArgs &A, const TaskRec *h_recs, int sm_count) {
  static thread_local uint32_t cached_smem = ~0u, cached_per_sm = 0;  // (per instantiation and host thread)
  if (cached_smem != smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, llk_flow_kernel<NPC>, 128, smem) != cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    cached_per_sm = (uint32_t)per_sm;
    cached_smem = smem;
  }
  static thread_local FlowArgs F;  // (31 KB: off the stack)
  F.n_quads = std::max(1u, A.n_bins_max / vb2::kBinsPerCta);
  F.stage_bytes = A.stage_bytes;
  F.pad_ = 0;
  memcpy(F.phred, A.phred, sizeof(F.phred));
  uint32_t launch = 0;
  for (uint32_t off = 0; off < A.n_jobs; off += (uint32_t)kFlowJobs, ++launch) {
    F.recs = A.recs + off;
    F.n_jobs = std::min<uint32_t>((uint32_t)kFlowJobs, A.n_jobs - off);
    F.queue = A.queue + launch;  // (one task counter per launch of the batch)
    for (uint32_t i = 0; i < F.n_jobs; ++i) {
      const JobParams &J = h_recs[off + i].J;
      FlowJob &D = F.jobs[i];
      for (int p = 0; p < kNumPairs; ++p) {  // the products llk_stream_kernel forms on the device (load_quad)
        D.C1[p] = J.c0[p] * J.c1[p];
        D.C2[p] = J.c1[p] * J.c1[p];
        D.c0[p] = J.c0[p];
        D.c1[p] = J.c1[p];
      }
      for (int k = 0; k < 4; ++k) {
        D.pc1[k] = J.pc1[k];
        D.pc2[k] = J.pc2[k];
      }
    }
    const uint32_t n_tasks = F.n_jobs * A.n_bins_max;
    const uint32_t grid = std::max(1u, std::min(cached_per_sm * (uint32_t)std::max(1, sm_count), (n_tasks + 3u) / 4u));
    // The launches of a batch do not depend on each other (own jobs, own task counter, own partial-sum slots): with
    // programmatic stream serialization the CTAs of launch i+1 move in as those of launch i run out of tasks, so the
    // cut into launches costs no drained-GPU gap (llk_reduce_kernel behind them is an ordinary launch: it waits for all).
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (launch > 0 && flow_overlap()) ? 1 : 0;  // (the first launch of a batch waits for whatever precedes the batch)
    cudaLaunchKernelEx(&cfg, llk_flow_kernel<NPC>, F);
  }
}
static_assert((VB2_MAX_BATCH + kFlowJobs - 1) / kFlowJobs <= (int)kQueueWords, "one task counter per launch of a batch");
// 0: let the shape decide; 1: always llk_stream_kernel (task queue); set by VB2_STREAM_KERNEL=queue (A/B runs, tests)
int stream_kernel_choice() {  // (read at every launch: the tests switch it inside one process)
  const char *e = getenv("VB2_STREAM_KERNEL");
  return (e && !strcmp(e, "queue")) ? 1 : 0;
}
// llk_flow_kernel serves the fixed layouts whose blobs fit a stage and whose bins have at most 32 rounds (one item table
// per task); everything else goes to llk_stream_kernel
bool flow_shape(int spec, bool chunked) { return !chunked && (spec == 2 || spec == 4) && stream_kernel_choice() == 0; }
void launch_stream(uint32_t smem, cudaStream_t stream, const LaunchArgs &A, const TaskRec *h_recs, int spec, bool chunked,
                   int sm_count) {
  bool flow = flow_shape(spec, chunked);
  for (uint32_t j = 0; flow && j < A.n_jobs; ++j) flow = h_recs[j].S.n_rounds <= 32u;
  if (flow) {
    if (spec == 2) launch_flow_as<2>(smem, stream, A, h_recs, sm_count);
    else launch_flow_as<4>(smem, stream, A, h_recs, sm_count);
  } else if (chunked) launch_stream_as<0, true>(smem, stream, A, sm_count);
  else if (spec == 2) launch_stream_as<2, false>(smem, stream, A, sm_count);
  else if (spec == 4) launch_stream_as<4, false>(smem, stream, A, sm_count);
  else launch_stream_as<0, false>(smem, stream, A, sm_count);
  llk_reduce_kernel<<<dim3(A.n_jobs, 1, 1), dim3(32, 1, 1), 0, stream>>>(A);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

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
  int spec = 0;          // 2 / 4: FixedLayout<spec> applies to this sample, else 0
  bool chunked = false;  // some blob is deeper than one shared-memory stage
  std::vector<vb2::Round> rounds;  // host copy (kernel arguments)
  Slot *h_mbox = nullptr, *d_mbox = nullptr;
  uint32_t mbox_slots = 0;
  // Staging of parameter sets that do not fit in the kernel arguments: TWO sets used alternately, each guarded by
  // an event recorded behind the launch that reads it, so staging launch i+1 never waits for launch i.
  JobParams *h_jobs2 = nullptr, *d_jobs2 = nullptr;  // [2][VB2_MAX_BATCH] pinned / device
  JobParams *h_jobs = nullptr, *d_jobs = nullptr;    // the set in use
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  bool stage_used[2] = {false, false};
  int stage_idx = 0;
  double *d_out = nullptr;      // [VB2_MAX_BATCH]
  unsigned int *d_queue = nullptr;  // llk_stream_kernel task queue
  unsigned long long *d_trace = nullptr;  // vb2_llk_trace: [grid_x][kTraceSlots]
  bool trace_on = false;
  // evaluation session (llk_session_kernel resident on the device)
  BellChunk *h_bell = nullptr, *d_bell = nullptr;  // host-mapped doorbell
  BellChunk *d_relay = nullptr;                    // its copy in HBM
  Slot *d_dmbox = nullptr;                         // per-CTA partials of a search on the device (HBM/L2), two banks
  MinRequest *d_minreq_fwd = nullptr;              // the request as the head CTA forwards it to the others
  MinRequest *h_minreq = nullptr, *d_minreq = nullptr;  // host-mapped request / result of vb2_llk_minimize
  MinResult *h_minres = nullptr, *d_minres = nullptr;
  bool session_active = false;
  uint32_t session_relaunches = 0;
  double clock_khz = 0.0, session_idle_ms = 200.0;
  // eval_many staging (owned by the leading context)
  TaskRec *h_recs2 = nullptr, *d_recs2 = nullptr, *h_recs = nullptr, *d_recs = nullptr;  // [2][VB2_MAX_BATCH] / in use
  uint32_t many_n = 0, many_grid_x = 0, many_kc = 1, many_buf_bytes = 0;  // last staged eval_many launch
  int many_spec = 0;
  bool many_chunked = false;
  unsigned long long seq = 0;
  int pending_n = 0;                 // evaluations launched by eval_begin and not yet collected
  bool pending_host_reduce = false;
  unsigned long long pending_seq = 0;
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

int session_launch(vb2_llk_ctx *ctx, unsigned long long first_seq);
void session_stop(vb2_llk_ctx *ctx);
uint32_t env_kc(const char *name, uint32_t dflt);
void fill_phred(LaunchArgs *A);

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

// Pinned + device staging for parameter sets that do not fit in the kernel arguments (allocated on first use).
int ensure_job_staging(vb2_llk_ctx *ctx) {
  if (ctx->h_jobs2) return VB2_OK;
  VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_jobs2, sizeof(JobParams) * 2 * VB2_MAX_BATCH, cudaHostAllocDefault));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_jobs2, sizeof(JobParams) * 2 * VB2_MAX_BATCH));
  VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_recs2, sizeof(TaskRec) * 2 * VB2_MAX_BATCH, cudaHostAllocDefault));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_recs2, sizeof(TaskRec) * 2 * VB2_MAX_BATCH));
  for (int i = 0; i < 2; ++i) VB2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
  return VB2_OK;
}
// Take the other staging set (waiting, if ever needed, for the launch that read it two stagings ago).
int acquire_staging(vb2_llk_ctx *ctx) {
  int rc = ensure_job_staging(ctx);
  if (rc) return rc;
  const int i = ctx->stage_idx ^= 1;
  if (ctx->stage_used[i]) VB2_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[i]));
  ctx->stage_used[i] = true;
  const size_t o = (size_t)i * VB2_MAX_BATCH;
  ctx->h_jobs = ctx->h_jobs2 + o; ctx->d_jobs = ctx->d_jobs2 + o;
  ctx->h_recs = ctx->h_recs2 + o; ctx->d_recs = ctx->d_recs2 + o;
  return VB2_OK;
}
// Behind every launch that reads the staging set in use.
void release_staging(vb2_llk_ctx *ctx) {
  if (ctx->stage_ev[ctx->stage_idx]) cudaEventRecord(ctx->stage_ev[ctx->stage_idx], ctx->stream);
}

// Device-side reduction scratch: one row of n_bins = 4*grid_x partials + one ticket per concurrent job.
int ensure_slots(vb2_llk_ctx *ctx, uint32_t need) {
  if (need <= ctx->slots) return VB2_OK;
  uint32_t n = ctx->slots ? ctx->slots : 8;
  while (n < need) n *= 2;
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double *partials = nullptr;
  unsigned int *tickets = nullptr;
  const size_t gx = ctx->S.grid_x ? ctx->S.grid_x : 1;
  VB2_CUDA(ctx, cudaMalloc(&partials, (size_t)n * gx * vb2::kBinsPerCta * sizeof(double)));
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

// Wait until mailbox slots [0, n_slots) all carry sequence number `seq`.
int wait_mailbox(vb2_llk_ctx *ctx, uint32_t n_slots, unsigned long long seq) {
  volatile Slot *mb = ctx->h_mbox;
  if (!ctx->spin) {
    VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n_slots; ++i)
      if (mb[i].seq != seq) return set_err(ctx, VB2_ERR_CUDA, "kernel finished without publishing its result");
    return VB2_OK;
  }
  // Poll the host-mapped sequence words: cheaper than a stream synchronise for a ~5 us kernel.
  unsigned long spins = 0;
  auto t0 = std::chrono::steady_clock::now();
  uint32_t i = 0;
  while (i < n_slots) {
    if (mb[i].seq == seq) {
      // the device's writes invalidated these lines in the host's caches: once the first slot is in, ask for all
      // the others at once instead of missing on them one after the other (4 slots per 64-byte line)
      if (i == 0)
        for (uint32_t j = 4; j < n_slots; j += 4) __builtin_prefetch((const void *)(ctx->h_mbox + j), 0, 0);
      ++i;
      continue;
    }
    if ((++spins & 0xFFFFu) == 0) {
      cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady)
        return set_err(ctx, VB2_ERR_CUDA, std::string("llk_kernel: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && mb[i].seq != seq) {
        if (ctx->session_active) {  // the resident kernel left on its idle watchdog: bring it back for this doorbell
          ++ctx->session_relaunches;
          int rc = session_launch(ctx, seq);
          if (rc) return rc;
          continue;
        }
        return set_err(ctx, VB2_ERR_CUDA, "kernel finished without publishing its result");
      }
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > ctx->spin_timeout_ms) return set_err(ctx, VB2_ERR_TIMEOUT, "timed out waiting for the device");
    }
  }
  return VB2_OK;
}

// ---- evaluation session (llk_session_kernel) ---------------------------------------------------------------
struct SessionGeometry {
  uint32_t kc, n_items, smem;
  bool ok;
};
SessionGeometry session_geometry(const vb2_llk_ctx *ctx) {
  SessionGeometry g{1, 1, 0, false};
  const SampleDev &S = ctx->S;
  if (S.grid_x == 0 || ctx->chunked || ctx->rounds.size() > (size_t)kMaxArgRounds) return g;
  uint32_t want = (uint32_t)kSessionMaxConc;
  if (const char *t = getenv("VB2_LLK_SESSION_KC")) want = (uint32_t)std::min<int>(std::max(1, atoi(t)), kSessionMaxConc);
  g.kc = std::max(1u, std::min(want, S.n_rounds));
  g.n_items = (S.n_rounds + g.kc - 1) / g.kc;
  g.smem = 4u * g.kc * g.n_items * S.buf_bytes + S.n_rounds * 1024u;
  g.ok = g.n_items <= (uint32_t)kMaxSessionItems && g.smem <= 200u * 1024u;
  return g;
}
// (re)launch the resident kernel; it serves sequence numbers from ctx->seq + 1 - pending on
int session_launch(vb2_llk_ctx *ctx, unsigned long long first_seq) {
  const SessionGeometry g = session_geometry(ctx);
  if (!g.ok) return set_err(ctx, VB2_ERR_INVALID, "this sample cannot be held resident in shared memory");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));  // (a re-launch from the wait loop may find another device current)
  SessionArgs A;
  memset(&A, 0, sizeof(A));
  A.sample = ctx->S;
  A.bell = ctx->d_bell;
  A.relay = ctx->d_relay;
  VB2_CUDA(ctx, cudaMemsetAsync(ctx->d_relay, 0, sizeof(BellChunk) * kMaxBellChunks, ctx->stream));
  A.mbox = ctx->d_mbox;
  A.dmbox = ctx->d_dmbox;
  VB2_CUDA(ctx, cudaMemsetAsync(ctx->d_dmbox, 0, sizeof(Slot) * 2 * (size_t)ctx->S.grid_x * ctx->S.grid_x, ctx->stream));
  A.min_req_dev = ctx->d_minreq_fwd;
  A.min_req = ctx->d_minreq;
  A.min_res = ctx->d_minres;
  A.first_seq = first_seq;
  A.idle_cycles = (unsigned long long)(ctx->session_idle_ms * ctx->clock_khz);
  A.kc = g.kc;
  A.n_items = g.n_items;
  A.n_chunks = 12u + 2u * ctx->S.n_pc;
  A.null_eval = getenv("VB2_LLK_SESSION_NULL") ? 1u : 0u;
  A.trace = ctx->trace_on ? ctx->d_trace : nullptr;
  LaunchArgs tmp;
  fill_phred(&tmp);
  memcpy(A.phred, tmp.phred, sizeof(A.phred));
  memcpy(A.rounds, ctx->rounds.data(), ctx->rounds.size() * sizeof(vb2::Round));
  const dim3 grid(ctx->S.grid_x, 1, 1), block(128u * g.kc, 1, 1);
  if (ctx->spec == 2) llk_session_kernel<2><<<grid, block, g.smem, ctx->stream>>>(A);
  else if (ctx->spec == 4) llk_session_kernel<4><<<grid, block, g.smem, ctx->stream>>>(A);
  else llk_session_kernel<0><<<grid, block, g.smem, ctx->stream>>>(A);
  VB2_CUDA(ctx, cudaGetLastError());
  return VB2_OK;
}
// ring the doorbell for sequence number `seq`: every chunk's payload first, its stamp after it (x86 keeps the
// order of the two stores; the device reads a chunk with one 16-byte load)
void session_ring(vb2_llk_ctx *ctx, const JobParams &J, unsigned long long seq, bool minimize = false) {
  const uint32_t k = ctx->S.n_pc;
  volatile BellChunk *bell = ctx->h_bell;
  unsigned long long xsum = 0ull;
  auto put = [&](uint32_t i, double v) {
    unsigned long long bits;
    memcpy(&bits, &v, sizeof(bits));
    xsum ^= bits;
    bell[i].payload = v;
    __atomic_store_n(&ctx->h_bell[i].seq, seq, __ATOMIC_RELEASE);
  };
  for (uint32_t p = 0; p < (uint32_t)kNumPairs; ++p) {
    put(p, J.c0[p]);
    put(kNumPairs + p, J.c1[p]);
  }
  for (uint32_t j = 0; j < k; ++j) {
    put(12u + j, J.pc1[j]);
    put(12u + k + j, J.pc2[j]);
  }
  // the control chunk: bits 8..63 of the payloads' checksum (a doorbell read torn between a new stamp and an old
  // payload cannot pass), the request type next to the sequence number
  const unsigned long long ctl = xsum & ~0xFFull;
  memcpy((void *)&ctx->h_bell[12u + 2u * k].payload, &ctl, sizeof(ctl));
  __atomic_store_n(&ctx->h_bell[12u + 2u * k].seq, seq | (minimize ? kCtlMinimize : 0ull), __ATOMIC_RELEASE);
}
// stop the resident kernel (exit doorbell) and wait for it; no-op without a session
void session_stop(vb2_llk_ctx *ctx) {
  if (!ctx || !ctx->session_active) return;
  __atomic_store_n(&ctx->h_bell[0].seq, kBellExit, __ATOMIC_RELEASE);
  cudaStreamSynchronize(ctx->stream);
  __atomic_store_n(&ctx->h_bell[0].seq, 0ull, __ATOMIC_RELEASE);
  ctx->session_active = false;
}

enum class Reduce { kHost, kDevice };

// Launch geometry.  One evaluation per launch: up to kMaxConcRounds rounds of a bin in flight at once
// (4 warps per round).  Several evaluations per launch: four-warp CTAs, four co-resident per SM, every warp
// walks all rounds of its bin, so one CTA's data wait and reduction tail hide behind the others' arithmetic.
struct Geometry {
  uint32_t kc, n_buf, threads, smem;
};
uint32_t env_kc(const char *name, uint32_t dflt) {
  if (const char *t = getenv(name)) return (uint32_t)std::min<int>(std::max(1, atoi(t)), kMaxConcRounds);
  return dflt;
}
constexpr uint32_t kMaxMeetRounds = 64;  // kc > 1 keeps n_rounds KiB of marginals in shared memory
Geometry geometry(const vb2_llk_ctx *ctx, bool throughput) {
  Geometry g;
  const SampleDev &S = ctx->S;
  const uint32_t want = throughput ? 1u : env_kc("VB2_LLK_LAT_KC", (uint32_t)kMaxConcRounds);
  g.kc = std::max(1u, std::min(want, S.n_rounds));
  if (S.n_rounds > kMaxMeetRounds) g.kc = 1;
  for (;;) {  // (wide blobs -- sixteen fp64 PCs -- with many rounds: fewer rounds in flight rather than no launch at all)
    g.n_buf = (g.kc < S.n_rounds || ctx->chunked) ? 2u : 1u;
    g.threads = 128u * g.kc;
    g.smem = 4u * g.kc * g.n_buf * S.buf_bytes + (g.kc > 1 ? S.n_rounds * 1024u : 0u);
    if (g.smem <= 200u * 1024u || g.kc == 1) break;
    --g.kc;
  }
  return g;
}

void fill_phred(LaunchArgs *A) {
  struct Table {
    double v[kPhredArgs];
    Table() {
      vb2::build_phred_table(v);
      for (int q = vb2::kNumQual; q < kPhredArgs; ++q) v[q] = 1.0;
    }
  };
  static const Table table;  // (thread-safe initialisation)
  memcpy(A->phred, table.v, sizeof(table.v));
}

// Job record of the many-evaluations kernel: sample, partial-sum slot, the head of the round table, parameters.
void fill_rec(TaskRec *R, const vb2_llk_ctx *c, uint32_t pslot, const double *pc1, const double *pc2, double alpha) {
  R->S = c->S;
  R->pslot = pslot;
  R->pad_ = 0;
  const size_t nr = std::min<size_t>(c->rounds.size(), (size_t)kRecRounds);
  memcpy(R->rounds, c->rounds.data(), nr * sizeof(vb2::Round));
  if (nr < (size_t)kRecRounds) memset(R->rounds + nr, 0, ((size_t)kRecRounds - nr) * sizeof(vb2::Round));
  fill_job(&R->J, c->S.n_pc, pc1, pc2, alpha);
}

// Launch n evaluations of ONE sample.
//   Reduce::kHost    n <= kMaxArgJobs: per-CTA partials go to mailbox slots [j*grid_x + cta]
//   Reduce::kDevice  the last CTA of job j writes d_out[j] (if given) and mailbox slot [j] (if to_mailbox)
int launch_batch(vb2_llk_ctx *ctx, int n, const double *pc1, const double *pc2, const double *alphas,
                 Reduce mode, double *d_out, bool to_mailbox, unsigned long long *seq_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (n <= 0 || n > VB2_MAX_BATCH) return set_err(ctx, VB2_ERR_INVALID, "batch size out of range");
  if (!pc1 || !pc2 || !alphas) return set_err(ctx, VB2_ERR_INVALID, "null parameter array");
  if (ctx->S.grid_x == 0) return VB2_OK;  // no usable marker: handled by the callers
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  session_stop(ctx);  // (any other launch on this context ends its evaluation session)
  const uint32_t k = ctx->S.n_pc;
  const bool args = n == 1 && ctx->rounds.size() <= (size_t)kMaxArgRounds;  // everything travels in the kernel arguments
  if (mode == Reduce::kHost && !args)
    return set_err(ctx, VB2_ERR_INVALID, "internal: host reduction is for one evaluation with ARGS");
  if (mode == Reduce::kDevice) {
    int rc = ensure_slots(ctx, (uint32_t)n);
    if (rc) return rc;
  }
  LaunchArgs A;
  memset(&A, 0, sizeof(A));  // (every pointer a kernel tests for null starts out null)
  A.sample = ctx->S;
  A.samples = nullptr; A.slots = nullptr; A.jobs_dev = nullptr; A.recs = nullptr;
  A.d_out = d_out;
  A.mbox = nullptr;
  A.seq = 0;
  A.n_jobs = (uint32_t)n;
  A.pad_ = 0;
  A.n_bins_max = vb2::kBinsPerCta * ctx->S.grid_x;
  A.stage_bytes = ctx->S.buf_bytes;
  A.queue = ctx->d_queue;
  A.trace = ctx->trace_on ? ctx->d_trace : nullptr;
  fill_phred(&A);
  const Geometry g = geometry(ctx, n > 1);
  A.kc = g.kc;
  A.n_buf = g.n_buf;
  if (args) {
    memcpy(A.rounds, ctx->rounds.data(), ctx->rounds.size() * sizeof(vb2::Round));
    fill_job(&A.jobs[0], k, pc1, pc2, alphas[0]);
  } else {
    int rcs = acquire_staging(ctx);
    if (rcs) return rcs;
    if (n == 1) {  // one evaluation whose round table does not fit in the arguments: parameters through HBM
      fill_job(&ctx->h_jobs[0], k, pc1, pc2, alphas[0]);
      VB2_CUDA(ctx, cudaMemcpyAsync(ctx->d_jobs, ctx->h_jobs, sizeof(JobParams), cudaMemcpyHostToDevice, ctx->stream));
      A.jobs_dev = ctx->d_jobs;
    } else {       // the many-evaluations kernel: one record per job
      for (int j = 0; j < n; ++j)
        fill_rec(&ctx->h_recs[j], ctx, (uint32_t)j, pc1 + (size_t)j * k, pc2 + (size_t)j * k, alphas[j]);
      VB2_CUDA(ctx, cudaMemcpyAsync(ctx->d_recs, ctx->h_recs, (size_t)n * sizeof(TaskRec), cudaMemcpyHostToDevice,
                                    ctx->stream));
      A.recs = ctx->d_recs;
    }
  }
  if (to_mailbox || mode == Reduce::kHost) {
    A.mbox = ctx->d_mbox;
    A.seq = ++ctx->seq;
    if (seq_out) *seq_out = A.seq;
  }
  if (n > 1) {  // several evaluations: the persistent task-queue kernel
    launch_stream(8u * A.stage_bytes, ctx->stream, A, ctx->h_recs, ctx->spec, ctx->chunked, ctx->sm_count);
    release_staging(ctx);
    VB2_CUDA(ctx, cudaGetLastError());
    return VB2_OK;
  }
  dim3 grid(ctx->S.grid_x, (unsigned)n, 1), block(g.threads, 1, 1);
  if (mode == Reduce::kHost) launch_llk<true, true>(grid, block, g.smem, ctx->stream, A, ctx->spec, ctx->chunked);
  else if (args) launch_llk<true, false>(grid, block, g.smem, ctx->stream, A, ctx->spec, ctx->chunked);
  else launch_llk<false, false>(grid, block, g.smem, ctx->stream, A, ctx->spec, ctx->chunked);
  if (!args) release_staging(ctx);
  VB2_CUDA(ctx, cudaGetLastError());
  return VB2_OK;
}

// Phred table + kernel attributes of the CURRENT device (idempotent; the stream may be the default one).
// Raise the dynamic shared-memory limit of the kernel instantiations a (layout, chunked) combination launches --
// only those: with lazy module loading every instantiation touched here is loaded, and there are 23 of them.
template <int NPC, bool CHUNKED>
cudaError_t raise_smem_limits(int bytes) {
  cudaError_t e = cudaFuncSetAttribute(llk_kernel<true, true, NPC, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_kernel<true, false, NPC, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_kernel<false, false, NPC, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_stream_kernel<NPC, CHUNKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_stream_kernel<NPC, CHUNKED>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if constexpr (!CHUNKED)
    if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_session_kernel<NPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if constexpr (!CHUNKED && NPC != 0) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_flow_kernel<NPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(llk_flow_kernel<NPC>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  }
  return e;
}
int init_device_tables(vb2_llk_ctx *ctx, int device, int spec, bool chunked) {
  static std::atomic<bool> done[64][4];  // [device][0: runtime layout, 1: NumPC 2, 2: NumPC 4, 3: chunked]
  const int which = chunked ? 3 : (spec == 2 ? 1 : spec == 4 ? 2 : 0);
  // (concurrent first calls -- the CLI's warm-up threads, one create per cohort worker -- may both set the
  // attributes: the calls are idempotent, the flag only saves the repeat)
  if (device >= 0 && device < 64 && done[device][which].load(std::memory_order_acquire)) return VB2_OK;
  const int smem_max = 200 * 1024;
  if (chunked) VB2_CUDA(ctx, (raise_smem_limits<0, true>(smem_max)));
  else if (spec == 2) VB2_CUDA(ctx, (raise_smem_limits<2, false>(smem_max)));
  else if (spec == 4) VB2_CUDA(ctx, (raise_smem_limits<4, false>(smem_max)));
  else VB2_CUDA(ctx, (raise_smem_limits<0, false>(smem_max)));
  if (device >= 0 && device < 64) done[device][which].store(true, std::memory_order_release);
  return VB2_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int vb2_abi_version(void) { return VB2_ABI_VERSION; }

int vb2_llk_batch_plan(const vb2_llk_ctx *ctx, int n, int *flow, int *kernel_launches, int *jobs_per_launch) {
  if (!ctx || n <= 0 || n > VB2_MAX_BATCH || !flow || !kernel_launches || !jobs_per_launch)
    return set_err(const_cast<vb2_llk_ctx *>(ctx), VB2_ERR_INVALID, "vb2_llk_batch_plan: bad argument");
  const bool f = n > 1 && flow_shape(ctx->spec, ctx->chunked) && ctx->S.n_rounds <= 32u;
  *flow = f ? 1 : 0;
  *kernel_launches = f ? (n + kFlowJobs - 1) / kFlowJobs : 1;
  *jobs_per_launch = f ? std::min(n, kFlowJobs) : n;
  return VB2_OK;
}

int vb2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vb2_llk_warmup(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(nullptr, VB2_ERR_NO_DEVICE, "no CUDA device available (this engine has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return set_err(nullptr, VB2_ERR_NO_DEVICE, "device out of range");
  VB2_CUDA(nullptr, cudaSetDevice(device));
  VB2_CUDA(nullptr, cudaFree(0));  // context creation
  return init_device_tables(nullptr, device, 2, false);  // module load of the common instantiations
}

const char *vb2_last_error(const vb2_llk_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

void vb2_llk_destroy(vb2_llk_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  session_stop(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (void *p : ctx->allocs) cudaFree(p);
  if (ctx->S.partials) cudaFree(ctx->S.partials);
  if (ctx->S.tickets) cudaFree(ctx->S.tickets);
  if (ctx->d_sample) cudaFree(ctx->d_sample);
  if (ctx->d_jobs2) cudaFree(ctx->d_jobs2);
  if (ctx->d_out) cudaFree(ctx->d_out);
  if (ctx->d_queue) cudaFree(ctx->d_queue);
  if (ctx->d_trace) cudaFree(ctx->d_trace);
  if (ctx->h_bell) cudaFreeHost(ctx->h_bell);
  if (ctx->d_relay) cudaFree(ctx->d_relay);
  if (ctx->d_dmbox) cudaFree(ctx->d_dmbox);
  if (ctx->d_minreq_fwd) cudaFree(ctx->d_minreq_fwd);
  if (ctx->h_minreq) cudaFreeHost(ctx->h_minreq);
  if (ctx->h_minres) cudaFreeHost(ctx->h_minres);
  if (ctx->d_recs2) cudaFree(ctx->d_recs2);
  if (ctx->h_mbox) cudaFreeHost(ctx->h_mbox);
  if (ctx->h_jobs2) cudaFreeHost(ctx->h_jobs2);
  if (ctx->h_recs2) cudaFreeHost(ctx->h_recs2);
  for (int i = 0; i < 2; ++i)
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

}  // extern "C"

namespace vb2 {
// ---- context construction, in two parts (shared by vb2_llk_create and the device ingest, llk_ingest.cu) --------------
// part 1: the device, the stream, the wait mode
int ctx_open(const CreateParams &cp, vb2_llk_ctx *ctx) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(ctx, VB2_ERR_NO_DEVICE, "no CUDA device available (this engine has no CPU fallback)");
  }
  if (cp.device < 0 || cp.device >= ndev) return set_err(ctx, VB2_ERR_NO_DEVICE, "desc.device out of range");
  ctx->device = cp.device;
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  VB2_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  if (prop.major < 10)
    return set_err(ctx, VB2_ERR_NO_DEVICE, std::string("device is not Blackwell (sm_100a) : ") + prop.name);
  ctx->sm_count = prop.multiProcessorCount;
  {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    ctx->clock_khz = khz > 0 ? (double)khz : 1.9e6;
    if (const char *t = getenv("VB2_LLK_SESSION_IDLE_MS")) ctx->session_idle_ms = std::max(1.0, atof(t));
  }
  if (cp.stream) {
    ctx->stream = static_cast<cudaStream_t>(cp.stream);
  } else {
    VB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  ctx->spin = !(cp.flags & VB2_FLAG_NO_SPIN);
  if (const char *t = getenv("VB2_LLK_SPIN_TIMEOUT_MS")) ctx->spin_timeout_ms = atof(t);
  if (cp.panel_dtype != VB2_PANEL_FP64 && cp.panel_dtype != VB2_PANEL_FP32)
    return set_err(ctx, VB2_ERR_INVALID, "unknown panel_dtype");
  return VB2_OK;
}
PackConfig ctx_pack_config(const vb2_llk_ctx *ctx, const CreateParams &cp) {
  PackConfig cfg;
  cfg.max_ctas = (uint32_t)ctx->sm_count;  // one persistent CTA per SM
  cfg.panel_fp64 = cp.panel_dtype == VB2_PANEL_FP64;
  if (cp.flags & VB2_FLAG_BATCHED) cfg.min_rounds = 5;
  if (const char *t = getenv("VB2_LLK_MAX_CTAS")) cfg.max_ctas = (uint32_t)std::max(1, atoi(t));
  return cfg;
}
cudaStream_t ctx_stream(const vb2_llk_ctx *ctx) { return ctx->stream; }
int ctx_error(vb2_llk_ctx *ctx, int code, const std::string &msg) { return set_err(ctx, code, msg); }

// part 2: adopt an image that already sits in device memory (d_blob belongs to the context from here on; P = its sizes,
// layout and round table)
int ctx_adopt_image(vb2_llk_ctx *ctx, const CreateParams &cp, const PackedSample &P, uint8_t *d_blob) {
  int rc = VB2_OK;
  ctx->meta = P;
  ctx->meta.blob = {};
  ctx->meta.marker_index = {};  // keep only the sizes
  ctx->rounds = P.rounds;
  if (d_blob) {
    ctx->allocs.push_back(d_blob);
    ctx->device_bytes += P.blob_bytes;
  }
  const bool panel_fp64 = cp.panel_dtype == VB2_PANEL_FP64;
  SampleDev &S = ctx->S;
  const vb2::Round *d_rounds = nullptr;
  if ((rc = upload(ctx, P.rounds, &d_rounds, false))) return rc;
  S.blob = d_blob;
  S.rounds = d_rounds;
  S.log_other_const = P.log_other_const;
  S.min_af = cp.min_af != 0.0 ? cp.min_af : 0.00005;  // h:94
  S.max_af = cp.max_af != 0.0 ? cp.max_af : 0.99995;  // h:95
  S.n_rounds = (uint32_t)P.rounds.size();
  S.n_bins = P.n_bins;
  S.grid_x = P.n_slices ? P.grid_x : 0u;
  S.conc_rounds = P.conc_rounds;
  S.n_pc = P.n_pc;
  S.panel_fp64 = panel_fp64 ? 1u : 0u;
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
  S.buf_bytes = S.off_words + S.chunk_rows * 128u + 128u;  // (+ one row: the read loop looks one row ahead)
  ctx->chunked = max_rows > S.chunk_rows;
  const bool default_clamps = S.min_af == 0.00005 && S.max_af == 0.99995;
  ctx->spec = (!panel_fp64 && !P.known_af && default_clamps && (P.n_pc == 2 || P.n_pc == 4)) ? (int)P.n_pc : 0;
  if (getenv("VB2_LLK_NO_SPEC")) ctx->spec = 0;  // (tests: force the runtime-layout kernel)
  if ((rc = init_device_tables(ctx, ctx->device, ctx->spec, ctx->chunked))) return rc;
  S.n_buf = 2u;
  {
    const Geometry g = geometry(ctx, false);  // the one-evaluation launch (the larger CTA of the two geometries)
    ctx->block_threads = g.threads;
    ctx->smem_bytes = g.smem;
  }
  if (ctx->smem_bytes > 200u * 1024u) return set_err(ctx, VB2_ERR_INVALID, "shared-memory stage too large (n_pc too big?)");

  // ---- result plumbing ---------------------------------------------------------------------------
  ctx->mbox_slots = std::max<uint32_t>(VB2_MAX_BATCH, kMaxArgJobs * std::max(1u, S.grid_x));
  VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_mbox, sizeof(Slot) * ctx->mbox_slots, cudaHostAllocMapped));
  memset(ctx->h_mbox, 0, sizeof(Slot) * ctx->mbox_slots);
  VB2_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->d_mbox, ctx->h_mbox, 0));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_out, sizeof(double) * VB2_MAX_BATCH));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_queue, kQueueWords * sizeof(unsigned int)));
  VB2_CUDA(ctx, cudaMemsetAsync(ctx->d_queue, 0, kQueueWords * sizeof(unsigned int), ctx->stream));
  VB2_CUDA(ctx, cudaMalloc(&ctx->d_sample, sizeof(SampleDev)));
  if ((rc = ensure_slots(ctx, 8))) return rc;
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VB2_OK;
}
vb2_llk_ctx *ctx_new() { return new (std::nothrow) vb2_llk_ctx(); }
// diagnostics: the image as it sits in device memory (tests compare it with the host flatten's)
int ctx_read_image(vb2_llk_ctx *ctx, uint8_t *dst, uint64_t n) {
  if (n > ctx->meta.blob_bytes) return set_err(ctx, VB2_ERR_INVALID, "image is smaller than that");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n) VB2_CUDA(ctx, cudaMemcpy(dst, ctx->S.blob, n, cudaMemcpyDeviceToHost));
  return VB2_OK;
}
}  // namespace vb2

extern "C" {

static int create_impl(const vb2_llk_desc *desc, vb2_llk_ctx *ctx) {
  const vb2::CreateParams cp{desc->device, desc->stream, desc->flags, desc->panel_dtype, desc->min_af, desc->max_af};
  int rc = vb2::ctx_open(cp, ctx);
  if (rc) return rc;
  // ---- flatten on the host ----------------------------------------------------------------------
  double phred[vb2::kNumQual];
  vb2::build_phred_table(phred);
  const vb2::PackConfig cfg = vb2::ctx_pack_config(ctx, cp);
  vb2::PackedSample P;
  std::string perr;
  rc = vb2::pack_sample(*desc, cfg, phred, &P, &perr);
  if (rc) return set_err(ctx, rc, perr);
  // ---- upload ---------------------------------------------------------------------------------
  uint8_t *d_blob = nullptr;
  if (!P.blob.empty()) {
    VB2_CUDA(ctx, cudaMalloc(&d_blob, P.blob.size()));
    cudaError_t e = cudaMemcpyAsync(d_blob, P.blob.data(), P.blob.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(d_blob);
      return set_err(ctx, VB2_ERR_CUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
  }
  return vb2::ctx_adopt_image(ctx, cp, P, d_blob);
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

static int begin_batch(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                       const double *alphas) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (ctx->pending_n) return set_err(ctx, VB2_ERR_INVALID, "an evaluation is already pending on this context");
  if (n <= 0 || n > VB2_MAX_BATCH) return set_err(ctx, VB2_ERR_INVALID, "batch size out of range");
  if (ctx->session_active && n == 1 && ctx->S.grid_x) {  // ring the resident kernel's doorbell: no launch
    JobParams J;
    fill_job(&J, ctx->S.n_pc, pc_contam, pc_intended, alphas[0]);
    const unsigned long long s = ++ctx->seq;
    session_ring(ctx, J, s);
    ctx->pending_n = 1;
    ctx->pending_host_reduce = true;
    ctx->pending_seq = s;
    return VB2_OK;
  }
  const bool host_reduce = n == 1 && ctx->rounds.size() <= (size_t)kMaxArgRounds;
  unsigned long long seq = 0;
  int rc = launch_batch(ctx, n, pc_contam, pc_intended, alphas, host_reduce ? Reduce::kHost : Reduce::kDevice,
                        nullptr, true, &seq);
  if (rc) return rc;
  ctx->pending_n = n;
  ctx->pending_host_reduce = host_reduce;
  ctx->pending_seq = seq;
  return VB2_OK;
}

static int end_batch(vb2_llk_ctx *ctx, double *llk_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (!llk_out) return set_err(ctx, VB2_ERR_INVALID, "null output pointer");
  const int n = ctx->pending_n;
  if (!n) return set_err(ctx, VB2_ERR_INVALID, "no evaluation pending on this context");
  ctx->pending_n = 0;
  const uint32_t gx = ctx->S.grid_x;
  if (gx == 0) {  // no usable marker: the reference's empty sum (h:231, :313)
    for (int j = 0; j < n; ++j) llk_out[j] = 0.0;
    return VB2_OK;
  }
  const bool host_reduce = ctx->pending_host_reduce;
  int rc = wait_mailbox(ctx, host_reduce ? (uint32_t)n * gx : (uint32_t)n, ctx->pending_seq);
  if (rc) return rc;
  volatile Slot *mb = ctx->h_mbox;
  if (host_reduce) {
    // Same fixed order as the device-side reduction (lane-strided sums, then a butterfly), so both
    // paths return identical bits for identical inputs.
    for (int j = 0; j < n; ++j) {
      double s[32], t[32];
      for (uint32_t l = 0; l < 32; ++l) {
        s[l] = 0.0;
        for (uint32_t c = l; c < gx; c += 32) s[l] += mb[(size_t)j * gx + c].val;
      }
      for (uint32_t o = 16; o; o >>= 1) {
        for (uint32_t l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ o];
        memcpy(s, t, sizeof(s));
      }
      llk_out[j] = s[0] + ctx->S.log_other_const;
    }
  } else {
    for (int j = 0; j < n; ++j) llk_out[j] = mb[j].val;
  }
  return VB2_OK;
}

int vb2_llk_eval_batch(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                       const double *alphas, double *llk_out) {
  if (ctx && !llk_out) return set_err(ctx, VB2_ERR_INVALID, "null output pointer");
  int rc = begin_batch(ctx, n, pc_contam, pc_intended, alphas);
  if (rc) return rc;
  return end_batch(ctx, llk_out);
}

int vb2_llk_eval(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha,
                 double *llk_out) {
  return vb2_llk_eval_batch(ctx, 1, pc_contam, pc_intended, &alpha, llk_out);
}

int vb2_llk_eval_begin(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha) {
  return begin_batch(ctx, 1, pc_contam, pc_intended, &alpha);
}

int vb2_llk_eval_end(vb2_llk_ctx *ctx, double *llk_out) { return end_batch(ctx, llk_out); }

int vb2_llk_eval_batch_device(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                              const double *alphas, double *d_llk_out) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (!d_llk_out) return set_err(ctx, VB2_ERR_INVALID, "null device output pointer");
  if (n <= 0 || n > VB2_MAX_BATCH) return set_err(ctx, VB2_ERR_INVALID, "batch size out of range");
  int rc = launch_batch(ctx, n, pc_contam, pc_intended, alphas, Reduce::kDevice, d_llk_out, false, nullptr);
  if (rc) return rc;
  if (ctx->S.grid_x == 0) VB2_CUDA(ctx, cudaMemsetAsync(d_llk_out, 0, sizeof(double) * n, ctx->stream));
  return VB2_OK;
}

// eval_many, step 1: stage the sample table, slots and parameters of an n-sample launch in HBM.
static int stage_many(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                      const double *alphas, bool *nothing_to_do) {
  if (!ctxs || n <= 0 || !ctxs[0]) return set_err(nullptr, VB2_ERR_INVALID, "null/empty context list");
  vb2_llk_ctx *lead = ctxs[0];
  if (n > VB2_MAX_BATCH) return set_err(lead, VB2_ERR_INVALID, "batch size out of range");
  if (!pc_contam || !pc_intended || !alphas) return set_err(lead, VB2_ERR_INVALID, "null argument");
  VB2_CUDA(lead, cudaSetDevice(lead->device));
  {
    int rcs = acquire_staging(lead);
    if (rcs) return rcs;
  }
  const uint32_t k = lead->S.n_pc;
  // slot of job j inside its sample = number of earlier jobs on the same context
  std::unordered_map<const vb2_llk_ctx *, uint32_t> seen;
  std::vector<uint32_t> slot_of((size_t)n);
  for (int j = 0; j < n; ++j) {
    vb2_llk_ctx *c = ctxs[j];
    if (!c) return set_err(lead, VB2_ERR_INVALID, "null context in list");
    session_stop(c);
    if (c->device != lead->device) return set_err(lead, VB2_ERR_INVALID, "contexts live on different devices");
    if (c->S.n_pc != k) return set_err(lead, VB2_ERR_INVALID, "contexts differ in n_pc");
    slot_of[j] = seen[c]++;  // (number of earlier jobs on the same context)
  }
  for (auto &kv : seen) {  // (before the records are filled: growing the slots moves S.partials)
    int rc = ensure_slots(const_cast<vb2_llk_ctx *>(kv.first), kv.second);
    if (rc) return set_err(lead, rc, kv.first->err);
  }
  bool any = false, chunked = false;
  int spec = ctxs[0]->spec;
  uint32_t grid_x = 0, kc = 1, buf_bytes = 0;
  for (int j = 0; j < n; ++j) {
    vb2_llk_ctx *c = ctxs[j];
    chunked = chunked || c->chunked;
    if (c->spec != spec) spec = 0;
    // the launch uses the largest geometry; every sample indexes shared memory with its own buf_bytes
    const Geometry g = geometry(c, true);
    grid_x = std::max(grid_x, c->S.grid_x);
    kc = std::max(kc, g.kc);
    buf_bytes = std::max(buf_bytes, c->S.buf_bytes);
    any |= c->S.grid_x > 0;
    fill_rec(&lead->h_recs[j], c, slot_of[j], pc_contam + (size_t)j * k, pc_intended + (size_t)j * k, alphas[j]);
  }
  *nothing_to_do = !any;
  lead->many_n = 0;
  if (!any) return VB2_OK;
  // (a sample without a usable marker has no active bin: llk_reduce_kernel returns the reference's empty sum, 0.0)
  VB2_CUDA(lead, cudaMemcpyAsync(lead->d_recs, lead->h_recs, sizeof(TaskRec) * n, cudaMemcpyHostToDevice, lead->stream));
  lead->many_n = (uint32_t)n;
  lead->many_grid_x = grid_x;
  lead->many_kc = kc;
  lead->many_buf_bytes = buf_bytes;
  lead->many_spec = spec;
  lead->many_chunked = chunked;
  return VB2_OK;
}

// eval_many, step 2: one launch over whatever stage_many staged last (device reduction).
struct vb2_peer {
  int device = 0;
  uint32_t rank = 0, world = 1;
  void *own = nullptr;                 // this rank's buffer (cudaMalloc, exported by IPC handle)
  void *opened[kMaxPeers] = {};        // the peers' buffers as mapped into this process
  PeerDev host{};                      // pointers into the buffers
  PeerDev *d_dev = nullptr;            // the same in device memory (llk_reduce_kernel reads it)
  unsigned long long seq = 0;          // launches so far (the same on every rank: they issue the same calls)
  bool connected = false;
  double clock_khz = 1.9e6;
};
static size_t peer_vals_bytes(uint32_t world) { return sizeof(double) * 2 * (size_t)world * VB2_MAX_BATCH; }

static int fire_many(vb2_llk_ctx *lead, bool to_mailbox, unsigned long long *seq_out, double *d_out = nullptr,
                     vb2_peer *peer = nullptr) {
  if (!lead->many_n) return set_err(lead, VB2_ERR_INVALID, "internal: nothing staged");
  LaunchArgs A;
  memset(&A, 0, sizeof(A));
  A.trace = nullptr;
#ifdef VB2_PHASE_CLOCK
  if (!lead->d_trace) {
    VB2_CUDA(lead, cudaMalloc(&lead->d_trace, sizeof(unsigned long long) * kTraceSlots * 1024));
    VB2_CUDA(lead, cudaMemset(lead->d_trace, 0, sizeof(unsigned long long) * kTraceSlots * 1024));
  }
  A.trace = lead->d_trace;
#endif
  fill_phred(&A);
  A.recs = lead->d_recs;
  A.n_jobs = lead->many_n;
  A.d_out = d_out ? d_out : lead->d_out;
  if (peer) {  // the shard sums go to the peers; the gather kernel behind writes d_out
    A.d_out = nullptr;
    A.peer = peer->d_dev;
    A.peer_seq = ++peer->seq;
  }
  A.kc = lead->many_kc;
  A.n_buf = 2;
  if (to_mailbox) {
    A.mbox = lead->d_mbox;
    A.seq = ++lead->seq;
    if (seq_out) *seq_out = A.seq;
  }
  A.n_bins_max = vb2::kBinsPerCta * lead->many_grid_x;
  A.stage_bytes = lead->many_buf_bytes;
  A.queue = lead->d_queue;
  {
    int rca = init_device_tables(lead, lead->device, lead->many_spec, lead->many_chunked);
    if (rca) return rca;
  }
  launch_stream(8u * A.stage_bytes, lead->stream, A, lead->h_recs, lead->many_spec, lead->many_chunked, lead->sm_count);
  if (peer) {
    const unsigned long long patience = (unsigned long long)(20000.0 * peer->clock_khz);  // 20 s
    llk_gather_kernel<<<dim3((lead->many_n + 255u) / 256u, 1, 1), dim3(256, 1, 1), 0, lead->stream>>>(
        peer->host, A.peer_seq, lead->many_n, d_out ? d_out : lead->d_out, patience);
  }
  release_staging(lead);
  VB2_CUDA(lead, cudaGetLastError());
  return VB2_OK;
}

// ---- peer buffers (marker shards, one process per GPU) ------------------------------------------------------------
int vb2_peer_create(int device, uint32_t rank, uint32_t world, vb2_peer **out, void *ipc_handle_out) {
  if (!out || !ipc_handle_out) return set_err(nullptr, VB2_ERR_INVALID, "null argument");
  *out = nullptr;
  if (world < 1 || world > (uint32_t)kMaxPeers || rank >= world) return set_err(nullptr, VB2_ERR_INVALID, "vb2_peer_create: bad rank / world (<= 8)");
  VB2_CUDA(nullptr, cudaSetDevice(device));
  vb2_peer *P = new (std::nothrow) vb2_peer();
  if (!P) return set_err(nullptr, VB2_ERR_NOMEM, "out of host memory");
  P->device = device; P->rank = rank; P->world = world;
  const size_t bytes = peer_vals_bytes(world) + sizeof(unsigned long long) * 2 * world + 64;
  cudaError_t e = cudaMalloc(&P->own, bytes);
  if (e == cudaSuccess) e = cudaMemset(P->own, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, P->own);
  if (e == cudaSuccess) e = cudaMalloc(&P->d_dev, sizeof(PeerDev));
  if (e != cudaSuccess) {
    if (P->own) cudaFree(P->own);
    delete P;
    return set_err(nullptr, VB2_ERR_CUDA, std::string("vb2_peer_create: ") + cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
  memcpy(ipc_handle_out, &h, sizeof(h));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  if (khz > 0) P->clock_khz = (double)khz;
  *out = P;
  return VB2_OK;
}

int vb2_peer_connect(vb2_peer *P, const void *ipc_handles) {
  if (!P || !ipc_handles) return set_err(nullptr, VB2_ERR_INVALID, "null argument");
  VB2_CUDA(nullptr, cudaSetDevice(P->device));
  for (uint32_t r = 0; r < P->world; ++r) {
    void *base = P->own;
    if (r != P->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const char *>(ipc_handles) + 64 * (size_t)r, sizeof(h));
      cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return set_err(nullptr, VB2_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      P->opened[r] = base;
    }
    P->host.vals[r] = static_cast<double *>(base);
    P->host.flags[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(base) + peer_vals_bytes(P->world));
  }
  P->host.ticket = reinterpret_cast<unsigned int *>(static_cast<char *>(P->own) + peer_vals_bytes(P->world) +
                                                     sizeof(unsigned long long) * 2 * P->world);
  P->host.world = P->world;
  P->host.rank = P->rank;
  VB2_CUDA(nullptr, cudaMemcpy(P->d_dev, &P->host, sizeof(PeerDev), cudaMemcpyHostToDevice));
  P->connected = true;
  return VB2_OK;
}

void vb2_peer_destroy(vb2_peer *P) {
  if (!P) return;
  cudaSetDevice(P->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < P->world; ++r)
    if (P->opened[r]) cudaIpcCloseMemHandle(P->opened[r]);
  if (P->own) cudaFree(P->own);
  if (P->d_dev) cudaFree(P->d_dev);
  delete P;
}

int vb2_llk_eval_many_device_peer(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                                  const double *alphas, vb2_peer *peer, double *d_llk_out) {
  if (!d_llk_out || !peer) return set_err(ctxs && n > 0 ? ctxs[0] : nullptr, VB2_ERR_INVALID, "null argument");
  if (!peer->connected) return set_err(ctxs && n > 0 ? ctxs[0] : nullptr, VB2_ERR_INVALID, "vb2_peer_connect has not been called");
  bool nothing = false;
  int rc = stage_many(ctxs, n, pc_contam, pc_intended, alphas, &nothing);
  if (rc) return rc;
  vb2_llk_ctx *lead = ctxs[0];
  if (nothing)  // (every rank must still take part in the exchange: a shard without a usable marker is a shard of zeros)
    return set_err(lead, VB2_ERR_INVALID, "vb2_llk_eval_many_device_peer: this rank's shard has no usable marker");
  return fire_many(lead, false, nullptr, d_llk_out, peer);
}

int vb2_llk_eval_many(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                      const double *alphas, double *llk_out) {
  if (!llk_out) return set_err(ctxs && n > 0 ? ctxs[0] : nullptr, VB2_ERR_INVALID, "null output pointer");
  unsigned long long seq = 0;
  bool nothing = false;
  int rc = stage_many(ctxs, n, pc_contam, pc_intended, alphas, &nothing);
  if (rc) return rc;
  if (nothing) {
    for (int j = 0; j < n; ++j) llk_out[j] = 0.0;
    return VB2_OK;
  }
  vb2_llk_ctx *lead = ctxs[0];
  if ((rc = fire_many(lead, true, &seq))) return rc;
  if ((rc = wait_mailbox(lead, (uint32_t)n, seq))) return rc;
  for (int j = 0; j < n; ++j) llk_out[j] = lead->h_mbox[j].val;
  return VB2_OK;
}

int vb2_llk_eval_many_device(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                             const double *alphas, double *d_llk_out) {
  if (!d_llk_out) return set_err(ctxs && n > 0 ? ctxs[0] : nullptr, VB2_ERR_INVALID, "null device output pointer");
  bool nothing = false;
  int rc = stage_many(ctxs, n, pc_contam, pc_intended, alphas, &nothing);
  if (rc) return rc;
  vb2_llk_ctx *lead = ctxs[0];
  if (nothing) {
    VB2_CUDA(lead, cudaMemsetAsync(d_llk_out, 0, sizeof(double) * n, lead->stream));
    return VB2_OK;
  }
  return fire_many(lead, false, nullptr, d_llk_out);
}

int vb2_llk_time_device_many(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup_launches, int launches,
                             const double *pc_contam, const double *pc_intended, double alpha, float *elapsed_ms) {
  if (!ctxs || n_ctx <= 0 || !ctxs[0] || launches <= 0 || !elapsed_ms) return set_err(nullptr, VB2_ERR_INVALID, "bad argument");
  vb2_llk_ctx *lead = ctxs[0];
  const uint32_t k = lead->S.n_pc;
  std::vector<double> pc1((size_t)n_ctx * k), pc2((size_t)n_ctx * k), al(n_ctx, alpha);
  for (int j = 0; j < n_ctx; ++j)
    for (uint32_t d = 0; d < k; ++d) {
      pc1[(size_t)j * k + d] = pc_contam[d] + (d == 0 ? 1e-7 * j : 0.0);  // a different point per sample
      pc2[(size_t)j * k + d] = pc_intended[d];
    }
  bool nothing = false;
  int rc = stage_many(ctxs, n_ctx, pc1.data(), pc2.data(), al.data(), &nothing);
  if (rc) return rc;
  if (nothing) return set_err(lead, VB2_ERR_INVALID, "no usable marker in any sample");
  cudaEvent_t e0, e1;
  VB2_CUDA(lead, cudaEventCreate(&e0));
  VB2_CUDA(lead, cudaEventCreate(&e1));
  for (int i = -warmup_launches; i < launches && rc == VB2_OK; ++i) {
    if (i == 0) cudaEventRecord(e0, lead->stream);
    rc = fire_many(lead, false, nullptr);
  }
  cudaEventRecord(e1, lead->stream);
  cudaError_t e = cudaEventSynchronize(e1);
  if (rc == VB2_OK && e != cudaSuccess) rc = set_err(lead, VB2_ERR_CUDA, cudaGetErrorString(e));
  if (rc == VB2_OK) cudaEventElapsedTime(elapsed_ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
#ifdef VB2_PHASE_CLOCK
  if (rc == VB2_OK && lead->d_trace) {
    unsigned long long ph[kTraceSlots];
    cudaMemcpy(ph, lead->d_trace, sizeof(ph), cudaMemcpyDeviceToHost);
    static const char *name[kTraceSlots] = {"loop head", "task switch", "mbar wait", "slice_begin", "ref run", "alt run",
                                            "finish", "produce", "#slices", "#tasks", "", "", "", "", "", ""};
    const double n_slices = (double)std::max(1ull, ph[8]);
    fprintf(stderr, "phase clock (cycles per slice per warp; %.0f slices, %.0f tasks):", n_slices, (double)ph[9]);
    for (int k = 0; k < 8; ++k) fprintf(stderr, " %s %.0f;", name[k], (double)ph[k] / n_slices);

    fprintf(stderr, "\n");
    cudaMemset(lead->d_trace, 0, sizeof(ph));
  }
#endif
  return rc;
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
    rc = launch_batch(c, 1, pc1.data(), pc_intended, &alpha, Reduce::kDevice, c->d_out, false, nullptr);
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
  double llk = 0.0, t_launch = 0.0, t_wait = 0.0;
  const bool breakdown = getenv("VB2_LLK_HOST_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0;
  for (int i = -warmup; i < steps; ++i) {
    if (i == 0) {
      VB2_CUDA(lead, cudaStreamSynchronize(lead->stream));
      t0 = std::chrono::steady_clock::now();
    }
    pc1[0] = pc_contam[0] + 1e-7 * ((i + warmup) % 1000);
    vb2_llk_ctx *c = ctxs[(i + warmup) % n_ctx];
    if (!breakdown) {
      int rc = vb2_llk_eval(c, pc1.data(), pc_intended, alpha, &llk);
      if (rc) return rc;
    } else {  // VB2_LLK_HOST_TIMING: the same two halves, timed separately
      auto a = std::chrono::steady_clock::now();
      int rc = vb2_llk_eval_begin(c, pc1.data(), pc_intended, alpha);
      auto b = std::chrono::steady_clock::now();
      if (rc == VB2_OK) rc = vb2_llk_eval_end(c, &llk);
      auto d = std::chrono::steady_clock::now();
      if (rc) return rc;
      if (i >= 0) {
        t_launch += std::chrono::duration<double>(b - a).count();
        t_wait += std::chrono::duration<double>(d - b).count();
      }
    }
  }
  *elapsed_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (breakdown)
    fprintf(stderr, "vb2_llk_time_host: %d steps, launch %.2f us + wait %.2f us per step (args %zu bytes)\n", steps,
            t_launch / steps * 1e6, t_wait / steps * 1e6, sizeof(LaunchArgs));
  if (last_llk) *last_llk = llk;
  return VB2_OK;
}

int vb2_llk_session_begin(vb2_llk_ctx *ctx) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (ctx->session_active) return VB2_OK;
  if (ctx->pending_n) return set_err(ctx, VB2_ERR_INVALID, "an evaluation is pending on this context");
  if (!ctx->spin) return set_err(ctx, VB2_ERR_INVALID, "a session needs the polling wait mode (VB2_FLAG_NO_SPIN is set)");
  if (!session_geometry(ctx).ok)
    return set_err(ctx, VB2_ERR_INVALID, "this sample cannot be held resident in shared memory (too deep or too large)");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->h_bell) {
    VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_bell, sizeof(BellChunk) * kMaxBellChunks, cudaHostAllocMapped));
    memset(ctx->h_bell, 0, sizeof(BellChunk) * kMaxBellChunks);
    VB2_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->d_bell, ctx->h_bell, 0));
    VB2_CUDA(ctx, cudaMalloc(&ctx->d_relay, sizeof(BellChunk) * kMaxBellChunks));
    VB2_CUDA(ctx, cudaMalloc(&ctx->d_dmbox, sizeof(Slot) * 2 * (size_t)std::max(1u, ctx->S.grid_x) * std::max(1u, ctx->S.grid_x)));
    VB2_CUDA(ctx, cudaMalloc(&ctx->d_minreq_fwd, sizeof(MinRequest)));
    VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_minreq, sizeof(MinRequest), cudaHostAllocMapped));
    VB2_CUDA(ctx, cudaHostAlloc((void **)&ctx->h_minres, sizeof(MinResult), cudaHostAllocMapped));
    memset(ctx->h_minreq, 0, sizeof(MinRequest));
    memset(ctx->h_minres, 0, sizeof(MinResult));
    VB2_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->d_minreq, ctx->h_minreq, 0));
    VB2_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->d_minres, ctx->h_minres, 0));
  }
  int rc = session_launch(ctx, ctx->seq + 1);
  if (rc) return rc;
  ctx->session_active = true;
  return VB2_OK;
}

int vb2_llk_session_end(vb2_llk_ctx *ctx) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (ctx->pending_n) return set_err(ctx, VB2_ERR_INVALID, "an evaluation is pending on this context");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  session_stop(ctx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(ctx, VB2_ERR_CUDA, std::string("llk_session_kernel: ") + cudaGetErrorString(e));
  return VB2_OK;
}

int vb2_llk_minimize(vb2_llk_ctx *ctx, const vb2_llk_model *model, const double *start, double scale, double ftol,
                     int64_t cycle_max, double llk1_in, vb2_llk_min_result *result) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  if (!model || !start || !result) return set_err(ctx, VB2_ERR_INVALID, "null argument");
  if (model->struct_size != sizeof(vb2_llk_model) || result->struct_size != sizeof(vb2_llk_min_result))
    return set_err(ctx, VB2_ERR_INVALID, "struct_size mismatch (ABI version skew)");
  const uint32_t k = ctx->S.n_pc;
  if (model->dim < 1 || model->dim > (uint32_t)kMinDim || k > (uint32_t)kMinPc)
    return set_err(ctx, VB2_ERR_INVALID, "vb2_llk_minimize: dim must be in [1, VB2_MIN_MAX_DIM] and n_pc <= 4");
  if (!ctx->session_active || ctx->S.grid_x == 0)
    return set_err(ctx, VB2_ERR_INVALID, "vb2_llk_minimize needs an open evaluation session on a sample with usable markers");
  if (ctx->pending_n) return set_err(ctx, VB2_ERR_INVALID, "an evaluation is pending on this context");
  for (uint32_t j = 0; j < k; ++j)
    if (model->pc1_from[j] >= (int32_t)model->dim || model->pc2_from[j] >= (int32_t)model->dim)
      return set_err(ctx, VB2_ERR_INVALID, "vb2_llk_model: index beyond dim");
  if (model->alpha_from >= (int32_t)model->dim) return set_err(ctx, VB2_ERR_INVALID, "vb2_llk_model: index beyond dim");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  MinRequest rq;
  memset(&rq, 0, sizeof(rq));
  rq.dim = (int32_t)model->dim;
  rq.n_pc = (int32_t)k;
  rq.alpha_from = model->alpha_from < 0 ? -1 : model->alpha_from;
  for (uint32_t j = 0; j < (uint32_t)kMinPc; ++j) {
    rq.pc1_from[j] = j < k && model->pc1_from[j] >= 0 ? model->pc1_from[j] : -1;
    rq.pc2_from[j] = j < k && model->pc2_from[j] >= 0 ? model->pc2_from[j] : -1;
    rq.pc1_fixed[j] = j < k ? model->pc1_fixed[j] : 0.0;
    rq.pc2_fixed[j] = j < k ? model->pc2_fixed[j] : 0.0;
  }
  rq.alpha_fixed = model->alpha_fixed;
  for (uint32_t j = 0; j < model->dim; ++j) rq.start[j] = start[j];
  rq.scale = scale;
  rq.ftol = ftol;
  rq.llk1 = llk1_in;
  rq.cycle_max = (long long)cycle_max;
  memcpy(ctx->h_minreq, &rq, sizeof(rq));
  __atomic_store_n(&ctx->h_minres->done, 0ull, __ATOMIC_RELEASE);
  const unsigned long long s0 = ctx->seq + 1;
  JobParams none;
  memset(&none, 0, sizeof(none));
  session_ring(ctx, none, s0, true);
  // wait for the result's stamp; a resident kernel that left on its idle watchdog is brought back (the bell still rings)
  unsigned long spins = 0;
  auto t0 = std::chrono::steady_clock::now();
  unsigned long long done = 0;
  while ((done = __atomic_load_n(&ctx->h_minres->done, __ATOMIC_ACQUIRE)) == 0ull) {
    if ((++spins & 0xFFFFu) == 0) {
      cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady)
        return set_err(ctx, VB2_ERR_CUDA, std::string("llk_session_kernel: ") + cudaGetErrorString(q));
      if (q == cudaSuccess && __atomic_load_n(&ctx->h_minres->done, __ATOMIC_ACQUIRE) == 0ull) {
        ++ctx->session_relaunches;
        int rc = session_launch(ctx, s0);
        if (rc) return rc;
      }
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > std::max(ctx->spin_timeout_ms, 120000.0)) return set_err(ctx, VB2_ERR_TIMEOUT, "timed out waiting for the device");
    }
  }
  ctx->seq = done;  // the search used the sequence numbers s0 .. done
  const MinResult &R = *ctx->h_minres;
  result->converged = R.converged;
  result->fmin = R.fmin;
  for (int j = 0; j < kMinDim; ++j) result->point[j] = R.point[j];
  result->evals = R.evals;
  result->cycle_count = R.cycle_count;
  result->llk1 = R.llk1;
  result->improved = R.improved;
  for (int j = 0; j < kMinPc; ++j) { result->best_pc_contam[j] = R.best_pc1[j]; result->best_pc_intended[j] = R.best_pc2[j]; }
  result->best_alpha = R.best_alpha;
  return VB2_OK;
}

int vb2_llk_trace(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha,
                  unsigned long long *stamps, uint32_t max_ctas, uint32_t *n_ctas, double *llk_out) {
  if (!ctx || !stamps || !n_ctas) return set_err(ctx, VB2_ERR_INVALID, "null argument");
  const uint32_t gx = ctx->S.grid_x;
  *n_ctas = gx;
  if (gx == 0 || gx > max_ctas) return set_err(ctx, VB2_ERR_INVALID, "stamp buffer too small (or no usable marker)");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  session_stop(ctx);
  const size_t bytes = (size_t)gx * kTraceSlots * sizeof(unsigned long long);
  if (!ctx->d_trace) VB2_CUDA(ctx, cudaMalloc(&ctx->d_trace, bytes));
  VB2_CUDA(ctx, cudaMemsetAsync(ctx->d_trace, 0, bytes, ctx->stream));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->trace_on = true;
  double llk = 0.0;
  int rc = VB2_OK;
  const char *trace_mode = getenv("VB2_LLK_TRACE_SESSION");
  if (trace_mode && !strcmp(trace_mode, "search")) {  // stages of the last evaluation of a search on the device
    rc = vb2_llk_session_begin(ctx);
    if (rc == VB2_OK) {
      const uint32_t k = ctx->S.n_pc;
      vb2_llk_model M;
      memset(&M, 0, sizeof(M));
      M.struct_size = sizeof(M);
      M.dim = 2 * k + 1;
      for (uint32_t j = 0; j < VB2_MAX_PC; ++j) { M.pc1_from[j] = j < k ? (int32_t)j : -1; M.pc2_from[j] = j < k ? (int32_t)(k + j) : -1; }
      M.alpha_from = (int32_t)(2 * k);
      std::vector<double> start(M.dim, 0.01);
      start[2 * k] = log(alpha / (1 - alpha));
      vb2_llk_min_result R;
      memset(&R, 0, sizeof(R));
      R.struct_size = sizeof(R);
      auto t0 = std::chrono::steady_clock::now();
      rc = vb2_llk_minimize(ctx, &M, start.data(), 1.0, 1e-8, 50000, 1e300, &R);
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      fprintf(stderr, "vb2_llk_trace: search on the device: %lld evaluations in %.1f us = %.2f us each\n", (long long)R.evals, us,
              R.evals ? us / (double)R.evals : 0.0);
      llk = -R.fmin;
    }
    if (rc == VB2_OK) rc = vb2_llk_session_end(ctx);
  } else if (trace_mode) {  // stages of the 20th evaluation of a session (see llk_session_kernel)
    rc = vb2_llk_session_begin(ctx);
    for (int i = 0; i < 20 && rc == VB2_OK; ++i) rc = vb2_llk_eval(ctx, pc_contam, pc_intended, alpha, &llk);
    if (rc == VB2_OK) rc = vb2_llk_session_end(ctx);
  } else {
    rc = vb2_llk_eval(ctx, pc_contam, pc_intended, alpha, &llk);
  }
  ctx->trace_on = false;
  if (rc) return rc;
  if (llk_out) *llk_out = llk;
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  VB2_CUDA(ctx, cudaMemcpy(stamps, ctx->d_trace, bytes, cudaMemcpyDeviceToHost));
  return VB2_OK;
}

int vb2_llk_sync(vb2_llk_ctx *ctx) {
  if (!ctx) return set_err(nullptr, VB2_ERR_INVALID, "null context");
  VB2_CUDA(ctx, cudaSetDevice(ctx->device));
  VB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VB2_OK;
}

}  // extern "C"
