/* vb2_llk.h -- C ABI of the B200 contamination-likelihood engine (libvb2llk.so).
 *
 * This is the drop-in boundary for ONE path of Griffan/VerifyBamID (VerifyBamID2):
 *
 *     double FullLLKFunc::ComputeMixLLKs(const std::vector<double>& tPC1,
 *                                        const std::vector<double>& tPC2, const double alpha)
 *                                                   (reference ContaminationEstimator.h:194-314)
 *
 * the genotype-mixture log-likelihood that FullLLKFunc::Evaluate (h:339-442), Initialize
 * (h:316-332) and CalculateLLK0 (h:334-337) call, hundreds of times per sample, from
 * AmoebaMinimizer::Minimize (MathGenMin.cpp:326-423).  The reference has no FFI; the line this
 * ABI cuts at is that one pure function plus the data it reads.  INTEGRATION.md shows the
 * five-line change a maintainer would make in ContaminationEstimator.h to bind it.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/CUDA/torch types cross the boundary;
 *   - every entry point returns VB2_OK (0) or a VB2_ERR_* code; the message is available from
 *     vb2_last_error().  No exceptions cross the boundary (the reference's error() throws,
 *     statgen/Error.cpp:26-40: the host wrapper converts);
 *   - a context is NOT re-entrant: evaluations on one context are issued from one host thread
 *     at a time, like the serial Nelder-Mead loop of the reference;
 *   - there is no CPU fallback: without a usable CUDA device every call fails with
 *     VB2_ERR_NO_DEVICE / VB2_ERR_CUDA.
 */
#ifndef VB2_LLK_H_
#define VB2_LLK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB2_ABI_VERSION 2
#define VB2_MAX_PC 16        /* largest n_pc a context accepts                               */
#define VB2_MAX_BATCH 4096   /* largest n in one vb2_llk_eval_batch / vb2_llk_eval_many call */

enum vb2_status {
  VB2_OK = 0,
  VB2_ERR_INVALID = 1,   /* bad argument / inconsistent descriptor                */
  VB2_ERR_NO_DEVICE = 2, /* no CUDA device, or desc.device out of range           */
  VB2_ERR_CUDA = 3,      /* a CUDA runtime call or kernel failed                  */
  VB2_ERR_NOMEM = 4,     /* host or device allocation failed                      */
  VB2_ERR_TIMEOUT = 5,   /* the device did not answer (VB2_LLK_SPIN_TIMEOUT_MS)   */
  VB2_ERR_UNSUPPORTED = 6 /* the device ingest does not take this input: use the host reader + vb2_llk_create */
};

enum vb2_panel_dtype {
  VB2_PANEL_FP32 = 0, /* UD and mu stored fp32 in HBM (north-star layout; LLK agrees with the
                         fp64 reference to ~1e-10 relative)                                  */
  VB2_PANEL_FP64 = 1  /* UD and mu stored fp64: AF = UD.PC + mu reproduces the reference's
                         double arithmetic operation for operation                           */
};

enum vb2_flags {
  VB2_FLAG_NO_SPIN = 1u << 0, /* wait with cudaStreamSynchronize instead of polling the
                                 host-mapped result mailbox                                  */
  VB2_FLAG_BATCHED = 1u << 1  /* the context serves batched evaluation (vb2_llk_eval_batch / _many, e.g. a marker
                                 shard of a multi-GPU run or a cohort member): lay a small sample out over fewer,
                                 deeper bins (>= 5 slices each) so that a task of the many-evaluations kernel
                                 amortises its set-up; one evaluation per launch then uses fewer SMs             */
};

typedef struct vb2_llk_ctx vb2_llk_ctx;

/* Everything ComputeMixLLKs reads, as flat arrays.  The nested vectors of the reference
 * become one CSR pair; all pointers are HOST pointers, copied at create (the caller may free
 * them afterwards).
 *
 *   reference object                                        field
 *   ------------------------------------------------------  ---------------------------------
 *   ContaminationEstimator::NumMarker          (h:446)      n_marker
 *   ContaminationEstimator::numPC              (h:49)       n_pc
 *   ContaminationEstimator::UD[i][k]           (h:452)      ud[i*ud_stride + k]
 *   ContaminationEstimator::means[i]           (h:454)      means[i]
 *   resolvedMarkers[i].baseInfoIndex           (h:471)      base_info_index[i]   (-1 = absent)
 *   resolvedMarkers[i].altBase                 (h:472)      alt_base[i]
 *   resolvedMarkers[i].knownAFValue, isAFknown (h:473,:44)  known_af[i] or NULL
 *   viewer.baseInfo[b] / viewer.qualInfo[b]    (SimplePileupViewer.h:94-95)
 *                                                           bases/quals[info_offset[b] ..
 *                                                                       info_offset[b+1])
 *   isSanityCheckDisabled                      (h:47)       sanity_disabled
 *   viewer.avgDepth, viewer.sdDepth            (SimplePileupViewer.h:100-101)
 *                                                           avg_depth, sd_depth
 *   FullLLKFunc::min_af, max_af                (h:94-95)    min_af, max_af (0,0 = defaults)
 */
typedef struct vb2_llk_desc {
  uint32_t struct_size; /* = sizeof(vb2_llk_desc); ABI guard */
  uint32_t n_marker;
  uint32_t n_pc;
  uint32_t ud_stride; /* doubles between consecutive rows of ud (>= n_pc) */
  const double *ud;
  const double *means;
  const int32_t *base_info_index;
  const char *alt_base;
  const double *known_af;
  const int64_t *info_offset;
  const char *bases;
  const char *quals;
  int32_t sanity_disabled;
  int32_t device;      /* CUDA device ordinal */
  double avg_depth;
  double sd_depth;
  double min_af;
  double max_af;
  int32_t panel_dtype; /* enum vb2_panel_dtype */
  uint32_t flags;      /* enum vb2_flags */
  /* Marker shard owned by this context (multi-GPU): the used markers are cut into 32-marker
   * slices and slice s belongs to shard s % shard_count.  0/0 and 0/1 mean "everything".
   * vb2_llk_eval then returns this shard's partial sum; the caller adds the partials
   * (one allreduce of a double per evaluation).                                             */
  uint32_t shard_rank;
  uint32_t shard_count;
  void *stream; /* cudaStream_t to launch on; NULL = a stream owned by the context */
  int64_t n_info; /* entries of info_offset minus one (viewer.baseInfo.size()); > 0: base_info_index is checked against
                     it, 0: the caller vouches for the indices                                                   */
} vb2_llk_desc;

typedef struct vb2_llk_info {
  uint32_t struct_size;
  uint32_t n_pc;
  uint64_t markers_used;     /* markers surviving the skip rules h:238-249 (this shard)     */
  uint64_t reads_used;       /* sum of their depths: R_used, the unit of the metric          */
  uint64_t reads_streamed;   /* ref+alt reads actually visited per evaluation                */
  uint64_t reads_folded;     /* class-2 ("other") reads folded into one constant at create   */
  uint64_t algorithmic_bytes;/* 2*R_used + 4*(n_pc+2)*markers_used              */
  uint64_t device_bytes;     /* bytes one evaluation really reads from HBM                   */
  uint32_t n_slices;         /* 32-marker warp slices                                        */
  uint32_t grid_x;           /* CTAs per evaluation                                          */
  uint32_t block_threads;
  uint32_t smem_bytes;       /* dynamic shared memory per CTA                                */
  int32_t device;
  int32_t sm_count;
  double log_other_const;    /* sum over folded reads of log(2e/3)                           */
} vb2_llk_info;

int vb2_abi_version(void);
int vb2_device_count(void); /* number of CUDA devices, 0 if none / no driver */

/* Optional: create the CUDA context and load the kernels of `device` now (0.3-0.5 s on a cold process).
 * Thread-safe; meant to run on a helper thread while the caller is still parsing its input files.     */
int vb2_llk_warmup(int device);

/* Flatten (classify, filter, sort, pack), upload to HBM, allocate the result mailbox.        */
int vb2_llk_create(const vb2_llk_desc *desc, vb2_llk_ctx **out);
void vb2_llk_destroy(vb2_llk_ctx *ctx);
int vb2_llk_get_info(const vb2_llk_ctx *ctx, vb2_llk_info *info);

/* One evaluation: *llk_out = ComputeMixLLKs(pc_contam, pc_intended, alpha)  (+LLK; the caller
 * negates, h:344).  Host pointers; returns after the result is in *llk_out.                  */
int vb2_llk_eval(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha,
                 double *llk_out);

/* The same evaluation split in two, so that one host thread can drive several contexts (marker shards
 * on several GPUs) concurrently: begin on every context, then end on every context and add the
 * partial sums in context order.  One evaluation may be pending per context.                    */
int vb2_llk_eval_begin(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha);
int vb2_llk_eval_end(vb2_llk_ctx *ctx, double *llk_out);

/* n evaluations of the same sample in ONE pass (e.g. all Nelder-Mead candidates of a step):
 * pc_contam / pc_intended are [n][n_pc] row-major, alphas and llk_out are [n].               */
int vb2_llk_eval_batch(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                       const double *alphas, double *llk_out);

/* Same as vb2_llk_eval_batch but asynchronous and with the results left in DEVICE memory
 * (d_llk_out[n], on the context's stream) -- the form a caller uses when the partial sums of
 * several marker shards go straight into an NCCL allreduce.                                  */
int vb2_llk_eval_batch_device(vb2_llk_ctx *ctx, int n, const double *pc_contam, const double *pc_intended,
                              const double *alphas, double *d_llk_out);

/* One evaluation of each of n DIFFERENT samples (contexts on the same device) in one launch:
 * job j uses ctxs[j] with parameters pc_contam[j][n_pc], pc_intended[j][n_pc], alphas[j].
 * All contexts must share n_pc.  Results and synchronisation follow ctxs[0]'s stream.        */
int vb2_llk_eval_many(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                      const double *alphas, double *llk_out);

/* Same, asynchronous, results left in DEVICE memory (d_llk_out[n], on ctxs[0]'s stream): the marker-shard
 * partial sums of n evaluations ready for ONE all-reduce.                                        */
int vb2_llk_eval_many_device(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                             const double *alphas, double *d_llk_out);

/* ---- the stage in front of the kernels, on the device: pileup text -> image ------------------------------------
 * For well-formed samtools-pileup text the device does what SimplePileupViewer::ReadPileup / ParsePileupSeqBasesOnly
 * (SimplePileupViewer.cpp:711-833), ContaminationEstimator::BuildResolvedMarkers (ContaminationEstimator.cpp:67-86) and
 * the host flatten of vb2_llk_create do, and leaves a context whose image is byte-identical to the one vb2_llk_create
 * builds from the host reader's arrays.  Three calls, with the caller's marker sanity check (IsSanityCheckOK,
 * cpp:543-587) between the last two:
 *   vb2_panel_create   the panel on the device, once for any number of samples: UD, mu, per row the chromosome (an index
 *                      into chrom_names), the 1-based position (PosVec) and the ALT base as resolved through ChooseBed;
 *   vb2_ingest_parse   text -> lines -> fields -> kept bases; join with the panel.  info->row_depth[i] = number of kept
 *                      bases on panel row i's pileup line, -1 when the pileup has no line for it; n_matched =
 *                      effectiveNumSite and num_bases = numBases as ReadPileup leaves them (cpp:826-831);
 *   vb2_ingest_flatten skip rules with the caller's avg_depth / sd_depth, class counts, marker order, slice geometry,
 *                      the fill -- all on the device; only the deal of ~3k slices to bins runs on the host.
 * Text that the reference parses with its stream-extraction quirks (short or empty lines, non-numeric position/depth,
 * duplicated positions, '.' reference allele next to '.'/',' bases, an indel length without digits, > 65,535 reads on a
 * site, >= 2 GiB of text) is answered with VB2_ERR_UNSUPPORTED: the caller then uses its host reader and vb2_llk_create,
 * which reproduce those quirks.                                                                                  */
typedef struct vb2_panel_desc {
  uint32_t struct_size, n_marker, n_pc, ud_stride;
  const double *ud, *means;
  const uint16_t *chrom_id;  /* per row: index into chrom_names                                                 */
  const int32_t *pos;        /* per row: 1-based position (column 3 of the .bed)                                */
  const char *alt_base;      /* per row: ChooseBed[chr][pos].second                                             */
  const char *chrom_names;   /* n_chrom NUL-terminated names, back to back                                      */
  uint32_t n_chrom;
  int32_t device;
} vb2_panel_desc;
typedef struct vb2_panel vb2_panel;
typedef struct vb2_ingest vb2_ingest;
typedef struct vb2_ingest_info {
  uint32_t struct_size, n_lines;
  uint32_t n_matched, pad_;
  uint64_t num_bases;
  const int32_t *row_depth;  /* [n_marker], owned by the ingest object                                          */
} vb2_ingest_info;
typedef struct vb2_flatten_desc {
  uint32_t struct_size;
  int32_t device;
  void *stream;              /* the context's stream (NULL: its own)                                            */
  uint32_t flags;            /* enum vb2_flags                                                                  */
  int32_t panel_dtype;       /* enum vb2_panel_dtype                                                            */
  double min_af, max_af;     /* 0, 0 = defaults                                                                 */
  int32_t sanity_disabled;
  uint32_t shard_rank, shard_count, pad_;  /* whole samples only: 0/0 or 0/1                                    */
  double avg_depth, sd_depth;
} vb2_flatten_desc;
int vb2_panel_create(const vb2_panel_desc *desc, vb2_panel **out);
void vb2_panel_destroy(vb2_panel *panel);
int vb2_ingest_parse(const vb2_panel *panel, const char *text, uint64_t n_bytes, vb2_ingest **out, vb2_ingest_info *info);
int vb2_ingest_flatten(vb2_ingest *ingest, const vb2_flatten_desc *desc, vb2_llk_ctx **out);
void vb2_ingest_destroy(vb2_ingest *ingest);
/* diagnostics: the first n_bytes of a context's image as it sits in device memory (tests compare it with the host flatten) */
int vb2_llk_debug_image(vb2_llk_ctx *ctx, void *dst, uint64_t n_bytes);

/* ---- marker shards on several GPUs, one process per GPU: the collective fused into the kernels ------------------
 * Instead of an NCCL all-reduce behind vb2_llk_eval_many_device, the reduce kernel PUSHES every evaluation's shard sum
 * into every rank's buffer over NVLink (peer stores) and a small gather kernel on every rank adds the shards in rank
 * order: vb2_llk_eval_many_device_peer leaves the SUMS OVER ALL SHARDS in d_llk_out, on ctxs[0]'s stream, identical
 * bits on every rank.  Set-up: every rank creates its buffer (vb2_peer_create returns a 64-byte CUDA IPC handle), the
 * ranks exchange the handles by whatever means they have (bench.py: one all_gather), then vb2_peer_connect maps the
 * peers' buffers.  Every rank must issue the same sequence of calls (a launch waits for all ranks' sums).  world <= 8. */
typedef struct vb2_peer vb2_peer;
int vb2_peer_create(int device, uint32_t rank, uint32_t world, vb2_peer **out, void *ipc_handle_out /* 64 bytes */);
int vb2_peer_connect(vb2_peer *peer, const void *ipc_handles /* [world][64] bytes, in rank order */);
void vb2_peer_destroy(vb2_peer *peer);
int vb2_llk_eval_many_device_peer(vb2_llk_ctx *const *ctxs, int n, const double *pc_contam, const double *pc_intended,
                                  const double *alphas, vb2_peer *peer, double *d_llk_out);

/* Evaluation session: for the several hundred DEPENDENT evaluations of one sample that a simplex search makes
 * (AmoebaMinimizer::Minimize, MathGenMin.cpp:326-423).  vb2_llk_session_begin launches one resident kernel that
 * keeps the whole sample in shared memory; until vb2_llk_session_end, vb2_llk_eval / _eval_begin / _eval_end on this
 * context ring a host-mapped doorbell instead of launching (same results, bit for bit, at about half the latency).
 * Rules: the sample must fit on chip (VB2_ERR_INVALID otherwise -- keep calling vb2_llk_eval without a session);
 * the resident kernel owns every SM of its device, so work of OTHER contexts on that device waits until the session
 * ends or has been idle for VB2_LLK_SESSION_IDLE_MS (default 200 ms; the kernel then leaves by itself and the next
 * evaluation brings it back); any batched call on this context ends the session.                              */
int vb2_llk_session_begin(vb2_llk_ctx *ctx);
int vb2_llk_session_end(vb2_llk_ctx *ctx);

/* ---- the simplex search next to the kernel ------------------------------------------------------------------
 * vb2_llk_minimize runs AmoebaMinimizer::Minimize (MathGenMin.cpp:326-423: Nelder-Mead, the reference's exact control
 * flow and operation order) ON THE DEVICE, inside the resident kernel of an evaluation session: the first warp of the
 * first CTA proposes the next point, every CTA evaluates its share of the markers from shared memory, the partial sums
 * meet in L2 -- no host round trip per evaluation; the host rings ONE doorbell and reads ONE result.
 *
 * The model says how a simplex vector v maps to ComputeMixLLKs' arguments -- the six branches of
 * FullLLKFunc::Evaluate (ContaminationEstimator.h:339-442) as a table:
 *   pc_contam[k]   = pc1_from[k] >= 0 ? v[pc1_from[k]] : pc1_fixed[k]
 *   pc_intended[k] = pc2_from[k] >= 0 ? v[pc2_from[k]] : pc2_fixed[k]
 *   alpha          = alpha_from  >= 0 ? InvLogit(v[alpha_from]) : alpha_fixed          (h:119-122)
 * and the objective is f(v) = -ComputeMixLLKs(pc_contam, pc_intended, alpha) (h:344).  Like Evaluate, the search keeps
 * the best point over ALL its evaluations: llk1 (in: the best value so far, out: the new one), and -- when it
 * improved -- the free components of the best point (best_pc_contam / best_pc_intended / best_alpha).
 * Limits: dim <= VB2_MIN_MAX_DIM, n_pc <= 4, an evaluation session must be open on the context (vb2_llk_session_begin);
 * otherwise VB2_ERR_INVALID -- the caller then drives the simplex itself with vb2_llk_eval, as the reference does.  */
#define VB2_MIN_MAX_DIM 9
typedef struct vb2_llk_model {
  uint32_t struct_size;
  uint32_t dim;                    /* length of the simplex vector                                          */
  int32_t pc1_from[VB2_MAX_PC];    /* index into v, or -1: fixed                                             */
  int32_t pc2_from[VB2_MAX_PC];
  int32_t alpha_from;
  int32_t pad_;
  double pc1_fixed[VB2_MAX_PC];
  double pc2_fixed[VB2_MAX_PC];
  double alpha_fixed;
} vb2_llk_model;

typedef struct vb2_llk_min_result {
  uint32_t struct_size;
  int32_t converged;               /* 0: cycleMax exceeded (Minimize returned numeric_limits<double>::max()) */
  double fmin;                     /* value at the best vertex                                               */
  double point[VB2_MIN_MAX_DIM];   /* the best vertex (AmoebaMinimizer::point on return)                     */
  int64_t evals;                   /* likelihood evaluations made                                            */
  int64_t cycle_count;             /* AmoebaMinimizer::cycleCount on return                                  */
  double llk1;                     /* best objective value over all evaluations, starting from llk1_in       */
  int32_t improved;                /* 1: some evaluation beat llk1_in; then the best_* fields are valid      */
  int32_t pad_;
  double best_pc_contam[4], best_pc_intended[4], best_alpha;
} vb2_llk_min_result;

/* start[dim]: the starting point; the initial simplex is start + scale * e_i (GeneralMinimizer::Reset, MathGenMin.cpp:17-25);
 * ftol, cycle_max: AmoebaMinimizer::Minimize's tolerance and cycleMax (50000 in the reference).                */
int vb2_llk_minimize(vb2_llk_ctx *ctx, const vb2_llk_model *model, const double *start, double scale, double ftol,
                     int64_t cycle_max, double llk1_in, vb2_llk_min_result *result);

/* How a batched call (vb2_llk_eval_batch / _many) over n evaluations of samples shaped like ctx's is launched:
 * *flow = 1 when it runs llk_flow_kernel (fp32 panel, NumPC 2 or 4, default AF clamps, no --KnownAF, every blob fits a
 * shared-memory stage; parameters in the kernel arguments), 0 for llk_stream_kernel (any shape); *kernel_launches = launches
 * of that kernel the batch is cut into (the one llk_reduce_kernel launch behind them not counted); *jobs_per_launch = the
 * most evaluations one of them carries.  Measurement aid (bench.py reports it); nothing is launched. */
int vb2_llk_batch_plan(const vb2_llk_ctx *ctx, int n, int *flow, int *kernel_launches, int *jobs_per_launch);

/* Block until everything queued on the context's stream has finished. */
int vb2_llk_sync(vb2_llk_ctx *ctx);

/* Message of the last failing call on this thread (ctx may be NULL); never NULL. */
const char *vb2_last_error(const vb2_llk_ctx *ctx);

/* ---- measurement helpers (used by bench.py; they launch exactly what vb2_llk_eval launches) ---
 * vb2_llk_time_device: `warmup` untimed then `steps` timed single-candidate evaluations issued
 * back to back from C, step i on ctxs[i % n_ctx] (all on ctxs[0]'s stream/device), bracketed by
 * CUDA events on that stream; results stay in device memory.  *elapsed_ms = the timed span.
 * vb2_llk_time_host: `steps` synchronous vb2_llk_eval calls (host parameter buffers in, host
 * scalar out, every step) timed with the host's steady clock; *elapsed_s = wall time.          */
int vb2_llk_time_device(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup, int steps, const double *pc_contam,
                        const double *pc_intended, double alpha, float *elapsed_ms);
/* vb2_llk_time_device_many: `launches` timed launches, each evaluating ALL n_ctx samples once (the
 * vb2_llk_eval_many kernel, results left in device memory): n_ctx steps per launch.             */
int vb2_llk_time_device_many(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup_launches, int launches,
                             const double *pc_contam, const double *pc_intended, double alpha, float *elapsed_ms);
int vb2_llk_time_host(vb2_llk_ctx *const *ctxs, int n_ctx, int warmup, int steps, const double *pc_contam,
                      const double *pc_intended, double alpha, double *elapsed_s, double *last_llk);

/* vb2_llk_trace: ONE vb2_llk_eval with the kernel's stage clock on.  stamps[cta][VB2_TRACE_SLOTS]:
 * SM clock (cycles) of the CTA's first thread at 0 entry, 1 tables published, 2 warp 0's last blob landed,
 * 3 warp 0 out of the read loop, 4 all marginals in shared memory, 5 CTA partial ready, 6 partial published;
 * slots 8 / 9 = the GPU's global timer (ns) at entry / exit; 7, 10..14 = finer points of the prologue and the
 * epilogue (see llk_kernel).  *n_ctas = CTAs of the launch (<= max_ctas).                                      */
#define VB2_TRACE_SLOTS 16
int vb2_llk_trace(vb2_llk_ctx *ctx, const double *pc_contam, const double *pc_intended, double alpha,
                  unsigned long long *stamps, uint32_t max_ctas, uint32_t *n_ctas, double *llk_out);

/* ---- diagnostics (host only, no CUDA call): the flattened image vb2_llk_create uploads ------
 * Lets CPU-only tests check the flatten (skip rules, classification, folding, slice layout,
 * sharding) without a GPU.  All pointers are owned by the view until vb2_llk_pack_free.     */
typedef struct vb2_packed_view {
  uint32_t struct_size;
  uint32_t n_pc, n_used, n_slices;
  uint32_t grid_x;       /* CTAs per launch: min(max_ctas, ceil(n_slices / 4))               */
  uint32_t n_bins;       /* 4 * grid_x: one bin per SM sub-partition                         */
  uint32_t conc_rounds;  /* rounds a CTA runs concurrently (it has 4 * conc_rounds warps)    */
  uint32_t n_rounds;     /* ceil(n_slices / n_bins)                                          */
  uint32_t max_stride;   /* largest blob stride in bytes                                     */
  uint32_t known_af;     /* 1: blobs carry known AF instead of UD/mu                         */
  uint32_t panel_elem;   /* 4 (fp32) or 8 (fp64): UD/mu element size inside the blobs        */
  uint32_t off_ud, off_mu, off_kaf, off_diag, off_words; /* byte offsets inside a blob       */
  uint64_t reads_used, reads_streamed, reads_folded, blob_bytes;
  double log_other_const;
  const uint8_t *blob;          /* [blob_bytes]; bin b's blob of round r at
                                   base_r + (b - first_bin_r) * stride_r;
                                   16-byte header = u32 ref_rows, alt_rows,
                                   n_valid | tail_ref << 8 | tail_alt << 12,
                                   full_ref_rows | full_alt_rows << 16                       */
  const uint32_t *rounds;       /* [n_rounds][6]: base lo, base hi, stride, first_bin, count, rows */
  const uint32_t *marker_index; /* [n_slices*32] panel row per (slice, lane), 0xFFFFFFFF pad;
                                   slice j (heaviest first) lives in round j / n_bins        */
  void *owner;
} vb2_packed_view;
/* max_ctas = SMs of the target device (0 = 148, a B200) */
int vb2_llk_pack_host(const vb2_llk_desc *desc, uint32_t max_ctas, vb2_packed_view *view);
void vb2_llk_pack_free(vb2_packed_view *view);

#ifdef __cplusplus
}
#endif
#endif /* VB2_LLK_H_ */
