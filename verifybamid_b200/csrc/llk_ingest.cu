// llk_ingest.cu -- the stage in front of the likelihood kernels, on the device: from the raw text of a samtools
// pileup to the image the kernels stream, without the host touching a read.  Replaces, for well-formed input,
//     SimplePileupViewer::ReadPileup / ParsePileupSeqBasesOnly      reference SimplePileupViewer.cpp:711-833
//     ContaminationEstimator::BuildResolvedMarkers                   reference ContaminationEstimator.cpp:67-86
//     the host flatten of llk_pack.cpp (skip rules h:238-249, classifyBase h:180-184, quality clamp h:296-298)
// in two steps with the host's marker sanity check (IsSanityCheckOK, cpp:543-587) between them:
//   vb2_ingest_parse    text -> lines -> fields -> kept bases per line; join with the panel's (chromosome, position)
//                       keys; per panel row: which line, how deep.  The host gets the depths (its sanity check and
//                       avgDepth / sdDepth are a few sums over 100k integers) and nothing else.
//   vb2_ingest_flatten  per-marker class counts, marker order, slice geometry (device) -> cost order, deal to bins,
//                       round table (host, llk_pack.cpp's plan_layout: 3k slices) -> blobs filled by one warp per
//                       slice straight from the text (device).  The image is byte-identical to the host flatten's.
// Text the reference parses with its stream-extraction quirks (short or empty lines, non-numeric fields, duplicated
// positions, a '.' reference allele next to '.'/',' bases, indel lengths without digits, more than 65,535 reads on a
// site) is NOT handled here: vb2_ingest_parse answers VB2_ERR_UNSUPPORTED and the caller runs the host reader, which
// reproduces those quirks (pileup_viewer.cpp).  Library code: cub::DeviceSelect / DeviceScan / DeviceRadixSort for
// the compactions and the two sorts (plumbing around the hand-written kernels).
#include <cuda_runtime.h>

#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "llk_internal.h"
#include "llk_pack.h"
#include "vb2_llk.h"

namespace {

using vb2::kNumQual;
constexpr uint32_t kAnomFields = 1u, kAnomInt = 2u, kAnomRefDot = 4u, kAnomQual = 8u, kAnomIndel = 16u, kAnomDup = 32u,
                   kAnomDeep = 64u;

thread_local std::string g_ingest_error;
int fail(int code, const std::string &msg) {
  g_ingest_error = msg;
  return vb2::ctx_error(nullptr, code, msg);
}
#define ING_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(e_ == cudaErrorMemoryAllocation ? VB2_ERR_NOMEM : VB2_ERR_CUDA,                   \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                              \
  } while (0)

struct LineRec {  // one pileup line: where its bases and qualities are, and its key
  uint32_t seq_off, seq_len, qual_off, qual_len;
  int32_t pos, chrom;  // chrom = -1: not a chromosome of the panel
};

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }
__device__ __forceinline__ bool is_digit(uint8_t c) { return c >= '0' && c <= '9'; }
__device__ __forceinline__ uint8_t to_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// ParsePileupSeqBasesOnly (SimplePileupViewer.cpp:711-746) as a visitor: f(base, quality char) for every kept base.
// Returns an anomaly mask (0 = parsed the way the reference parses it).
template <typename F>
__device__ __forceinline__ uint32_t for_each_kept_base(const uint8_t *seq, uint32_t n_seq, const uint8_t *qual, uint32_t n_qual, F f) {
  uint32_t iq = 0;
  for (uint32_t i = 0; i < n_seq; ++i) {
    const uint8_t c = seq[i];
    if (c == '+' || c == '-') {  // an indel: "+<n><n bases>", skipped, owns no quality (cpp:716-723)
      uint32_t t = i + 1, len = 0, digits = 0;
      while (t < n_seq && is_digit(seq[t])) {
        len = len * 10u + (uint32_t)(seq[t] - '0');
        ++t; ++digits;
      }
      if (digits == 0 || digits > 9) return kAnomIndel;  // (std::stoi would throw)
      i += digits + len;
    } else if (c == '^') {
      i += 1;  // read start: the next char is a mapping quality (cpp:724-726)
    } else if (c == '.' || c == ',' || c == 'A' || c == 'G' || c == 'C' || c == 'T' || c == 'N' || c == 'a' || c == 'g' ||
               c == 'c' || c == 't' || c == 'n') {
      if (iq >= n_qual) return kAnomQual;  // (the reference throws: fewer qualities than bases)
      f(c, qual[iq]);
      ++iq;
    } else if (c == '*' || c == '#') {
      ++iq;  // deletion placeholder: not modelled, but it owns a quality (cpp:738-743)
    }
  }
  return 0u;
}

struct IsNewline {
  const uint8_t *text;
  __device__ bool operator()(uint32_t i) const { return text[i] == '\n'; }
};

// One thread per line: the six fields `ss >> pChr >> pPos >> refAllele >> depth >> seq >> qual` reads (cpp:762-768),
// the kept-base count, and whether the line's position is one of the panel's.
__global__ void parse_lines_kernel(const uint8_t *text, uint64_t n_bytes, const uint32_t *newline, uint32_t n_newline, uint32_t n_lines,
                                   const char *names, const uint32_t *name_off, uint32_t n_chrom, const uint64_t *bed_keys,
                                   uint32_t n_bed, LineRec *recs, int32_t *depth, uint32_t *matched, uint32_t *anomaly,
                                   unsigned long long *num_bases) {
  const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  uint32_t p = line == 0 ? 0u : newline[line - 1] + 1u;
  const uint32_t end = line < n_newline ? newline[line] : (uint32_t)n_bytes;
  uint32_t tb[6], te[6];
  for (int k = 0; k < 6; ++k) {
    while (p < end && is_space(text[p])) ++p;
    tb[k] = p;
    while (p < end && !is_space(text[p])) ++p;
    te[k] = p;
    if (te[k] == tb[k]) {  // fewer than six fields: the reference would carry values over from the previous line
      atomicOr(anomaly, kAnomFields);
      matched[line] = 0;
      depth[line] = 0;
      return;
    }
  }
  auto parse_int = [&](int k, int32_t &out) {  // operator>>(int&): optional sign, digits; anything else is a quirk
    uint32_t q = tb[k];
    bool neg = false;
    if (text[q] == '+' || text[q] == '-') { neg = text[q] == '-'; ++q; }
    if (q == te[k] || te[k] - q > 9) return false;
    int32_t v = 0;
    for (; q < te[k]; ++q) {
      if (!is_digit(text[q])) return false;
      v = v * 10 + (text[q] - '0');
    }
    out = neg ? -v : v;
    return true;
  };
  int32_t pos = 0, dummy = 0;
  if (!parse_int(1, pos) || !parse_int(3, dummy)) {
    atomicOr(anomaly, kAnomInt);
    matched[line] = 0;
    depth[line] = 0;
    return;
  }
  LineRec R;
  R.seq_off = tb[4]; R.seq_len = te[4] - tb[4];
  R.qual_off = tb[5]; R.qual_len = te[5] - tb[5];
  R.pos = pos;
  R.chrom = -1;
  for (uint32_t c = 0; c < n_chrom; ++c) {
    const uint32_t len = name_off[c + 1] - name_off[c] - 1u;  // (names are NUL-terminated)
    if (len != te[0] - tb[0]) continue;
    bool same = true;
    for (uint32_t j = 0; j < len && same; ++j) same = (uint8_t)names[name_off[c] + j] == text[tb[0] + j];
    if (same) { R.chrom = (int32_t)c; break; }
  }
  // cpp:771-778: '.'/',' bases need a known reference allele
  const bool ref_is_dot = te[2] - tb[2] == 1 && text[tb[2]] == '.';
  uint32_t kept = 0;
  bool has_ref_base = false;
  uint32_t bad = for_each_kept_base(text + R.seq_off, R.seq_len, text + R.qual_off, R.qual_len,
                                    [&](uint8_t, uint8_t) { ++kept; });
  if (ref_is_dot)
    for (uint32_t j = 0; j < R.seq_len && !has_ref_base; ++j) has_ref_base = text[R.seq_off + j] == '.' || text[R.seq_off + j] == ',';
  if (has_ref_base) bad |= kAnomRefDot;
  if (kept > 65535u) bad |= kAnomDeep;
  if (bad) atomicOr(anomaly, bad);
  recs[line] = R;
  depth[line] = (int32_t)kept;
  uint32_t m = 0;
  if (R.chrom >= 0) {  // is (chromosome, position) in the panel's .bed?  (cpp:786-788)
    const uint64_t key = ((uint64_t)(uint32_t)R.chrom << 32) | (uint32_t)pos;
    uint32_t lo = 0, hi = n_bed;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (bed_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    m = lo < n_bed && bed_keys[lo] == key;
  }
  matched[line] = m;
  if (m) atomicAdd(num_bases, (unsigned long long)kept);  // numBases (cpp:826)
}

// matched line -> its info index (file order) and its key
__global__ void scatter_info_kernel(const uint32_t *matched, const uint32_t *rank, const LineRec *recs, uint32_t n_lines,
                                    uint32_t *info_line, uint64_t *info_key, uint32_t *info_id) {
  const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines || !matched[line]) return;
  const uint32_t g = rank[line];
  info_line[g] = line;
  info_key[g] = ((uint64_t)(uint32_t)recs[line].chrom << 32) | (uint32_t)recs[line].pos;
  info_id[g] = g;
}
__global__ void dup_check_kernel(const uint64_t *sorted_key, uint32_t n, uint32_t *anomaly) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && sorted_key[i] == sorted_key[i + 1]) atomicOr(anomaly, kAnomDup);  // cpp:812-824: the host reader's case
}
// BuildResolvedMarkers: panel row -> info index (or -1) and its depth (or -1)
__global__ void join_rows_kernel(const uint64_t *row_key, uint32_t n_rows, const uint64_t *sorted_key, const uint32_t *sorted_id,
                                 uint32_t n_info, const uint32_t *info_line, const int32_t *line_depth, int32_t *row_info,
                                 int32_t *row_depth) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const uint64_t key = row_key[i];
  uint32_t lo = 0, hi = n_info;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sorted_key[mid] < key) lo = mid + 1; else hi = mid;
  }
  int32_t g = -1, d = -1;
  if (lo < n_info && sorted_key[lo] == key) {
    g = (int32_t)sorted_id[lo];
    d = line_depth[info_line[g]];
  }
  row_info[i] = g;
  row_depth[i] = d;
}

// ---- flatten -------------------------------------------------------------------------------------------------------
// step 1 of llk_pack.cpp: skip rules (h:238-249) and per-marker class counts; class-2 reads go to a histogram
__global__ void marker_counts_kernel(const uint8_t *text, const LineRec *recs, const uint32_t *info_line, const int32_t *row_info,
                                     const int32_t *row_depth, const uint8_t *alt_base, uint32_t n_rows, int sanity_disabled,
                                     double lo_depth, double hi_depth, uint32_t *used, uint32_t *n_ref, uint32_t *n_alt,
                                     unsigned long long *other_hist, unsigned long long *totals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  uint32_t u = 0, nr = 0, na = 0, no = 0;
  const int32_t g = row_info[i], size = row_depth[i];
  if (g >= 0 && size > 0 && (sanity_disabled || !((double)size < lo_depth || (double)size > hi_depth))) {
    u = 1;
    const LineRec R = recs[info_line[g]];
    const uint8_t alt = to_upper(alt_base[i]);
    for_each_kept_base(text + R.seq_off, R.seq_len, text + R.qual_off, R.qual_len, [&](uint8_t b, uint8_t qc) {
      if (b == '.' || b == ',') ++nr;
      else if (to_upper(b) == alt) ++na;
      else {
        int q = (int)qc - 33;
        q = q < 0 ? 0 : (q > 93 ? 93 : q);
        atomicAdd(other_hist + q, 1ull);
        ++no;
      }
    });
    atomicAdd(totals + 0, (unsigned long long)size);       // reads_used
    atomicAdd(totals + 1, (unsigned long long)(nr + na));  // reads_streamed
    atomicAdd(totals + 2, (unsigned long long)no);         // reads_folded
  }
  used[i] = u;
  n_ref[i] = nr;
  n_alt[i] = na;
}
// step 2: the markers' sort key (see llk_pack.cpp), in panel order of the used markers
__global__ void marker_keys_kernel(const uint32_t *used, const uint32_t *rank, const uint32_t *n_ref, const uint32_t *n_alt,
                                   uint32_t n_rows, uint32_t *used_row, uint64_t *key, uint32_t *id) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows || !used[i]) return;
  const uint32_t u = rank[i];
  const uint64_t na = n_alt[i] < 0xFFFFFFu ? n_alt[i] : 0xFFFFFFu, nr = n_ref[i] < 0xFFFFFFu ? n_ref[i] : 0xFFFFFFu;
  used_row[u] = i;
  key[u] = ((0xFFFFFFull - na) << 24) | ((na & 1u) ? nr : 0xFFFFFFull - nr);
  id[u] = u;
}
// step 3a: geometry of every 32-marker slice of the sorted order
__global__ void slice_geom_kernel(const uint32_t *order, const uint32_t *used_row, const uint32_t *n_ref, const uint32_t *n_alt,
                                  uint32_t n_used, uint32_t n_slices, vb2::SliceGeom *geom) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slices) return;
  uint32_t wr = 0, wa = 0, fr = 0xFFFFFFFFu, fa = 0xFFFFFFFFu, nr0 = 0, na0 = 0;
  bool same_r = true, same_a = true;
  for (uint32_t l = 0; l < 32u; ++l) {
    const uint32_t o = s * 32u + l;
    if (o >= n_used) break;
    const uint32_t row = used_row[order[o]];
    const uint32_t nr = n_ref[row], na = n_alt[row];
    wr = max(wr, (nr + 3u) / 4u);
    wa = max(wa, (na + 3u) / 4u);
    fr = min(fr, nr / 4u);
    fa = min(fa, na / 4u);
    if (l == 0) { nr0 = nr; na0 = na; }
    same_r = same_r && nr == nr0;
    same_a = same_a && na == na0;
  }
  const uint32_t rr = wr - fr, ra = wa - fa;
  const bool tail_r = same_r && rr == 1, tail_a = same_a && ra == 1;
  vb2::SliceGeom G;
  G.wr = wr; G.wa = wa;
  G.cost = 4u * (fr + fa) + (tail_r ? 2u : 7u * rr) + (tail_a ? 2u : 7u * ra) + 15u;  // = vb2::slice_cost
  G.first = s * 32u;
  geom[s] = G;
}

struct BlobJob {  // one blob of the image: where it goes and which markers it holds
  uint64_t offset;
  uint32_t first, wr, wa, pad_;
};
struct FillTables {
  double a_ref[kNumQual][3], a_alt[kNumQual][3];
};
constexpr int kFillWarps = 2;
// step 6: one warp per blob, one lane per marker -- panel columns, diagonal products, quality bytes in ascending order
template <bool PANEL_FP64>
__global__ void __launch_bounds__(32 * kFillWarps) fill_blobs_kernel(
    const uint8_t *text, const LineRec *recs, const uint32_t *info_line, const int32_t *row_info, const uint8_t *alt_base,
    const double *ud, uint32_t ud_stride, const double *means, const uint32_t *order, const uint32_t *used_row, uint32_t n_used,
    const BlobJob *jobs, uint32_t n_blobs, uint32_t n_pc, vb2::BlobLayout L, const FillTables *tab, uint8_t *image) {
  __shared__ uint16_t s_hist[kFillWarps][2][kNumQual][32];  // [class][quality][lane]: conflict-free per-lane counters
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * kFillWarps + warp;
  if (q >= n_blobs) return;
  const BlobJob J = jobs[q];
  uint8_t *blob = image + J.offset;
  const uint32_t o = J.first + (uint32_t)lane;
  const bool valid = o < n_used;
  for (int c = 0; c < 2; ++c)
    for (int k = 0; k < kNumQual; ++k) s_hist[warp][c][k][lane] = 0;
  uint32_t nr = 0, na = 0, row = 0;
  LineRec R{};
  if (valid) {
    row = used_row[order[o]];
    R = recs[info_line[row_info[row]]];
    const uint8_t alt = to_upper(alt_base[row]);
    for_each_kept_base(text + R.seq_off, R.seq_len, text + R.qual_off, R.qual_len, [&](uint8_t b, uint8_t qc) {
      int k = (int)qc - 33;
      k = k < 0 ? 0 : (k > 93 ? 93 : k);
      if (b == '.' || b == ',') { ++s_hist[warp][0][k][lane]; ++nr; }
      else if (to_upper(b) == alt) { ++s_hist[warp][1][k][lane]; ++na; }
    });
  }
  // panel columns (neutral values in the padding lanes: UD 0, mu 1), as llk_pack.cpp writes them
  for (uint32_t k = 0; k <= n_pc; ++k) {  // k == n_pc: mu
    const double v = valid ? (k < n_pc ? ud[(size_t)row * ud_stride + k] : means[row]) : (k < n_pc ? 0.0 : 1.0);
    uint8_t *dst = blob + (k < n_pc ? L.off_ud + k * 32u * L.panel_elem : L.off_mu) + (uint32_t)lane * L.panel_elem;
    if (PANEL_FP64) *reinterpret_cast<double *>(dst) = v;
    else *reinterpret_cast<float *>(dst) = (float)v;
  }
  // the reads, class by class in ascending quality, four to a word; the diagonal products follow the same order
  double dg[3] = {1.0, 1.0, 1.0};
  uint32_t *words = reinterpret_cast<uint32_t *>(blob + L.off_words);
  for (int c = 0; c < 2; ++c) {
    const uint32_t t0 = c == 0 ? 0u : J.wr;
    uint32_t r = 0, word = 0xFFFFFFFFu;
    for (int k = 0; k < kNumQual; ++k) {
      const double a0 = c == 0 ? tab->a_ref[k][0] : tab->a_alt[k][0], a1 = c == 0 ? tab->a_ref[k][1] : tab->a_alt[k][1],
                   a2 = c == 0 ? tab->a_ref[k][2] : tab->a_alt[k][2];
      for (uint32_t n = s_hist[warp][c][k][lane]; n; --n, ++r) {
        dg[0] = __dmul_rn(dg[0], a0); dg[1] = __dmul_rn(dg[1], a1); dg[2] = __dmul_rn(dg[2], a2);
        const uint32_t sh = (r & 3u) * 8u;  // little-endian: byte b of the word = bits 8b..8b+7
        word = (word & ~(0xFFu << sh)) | ((uint32_t)k << sh);
        if ((r & 3u) == 3u) {
          words[(size_t)(t0 + r / 4u) * 32u + lane] = word;
          word = 0xFFFFFFFFu;
        }
      }
    }
    if (r & 3u) words[(size_t)(t0 + r / 4u) * 32u + lane] = word;  // the ragged last word keeps its 0xFF fillers
  }
  double *diag = reinterpret_cast<double *>(blob + L.off_diag);
  for (int g = 0; g < 3; ++g) diag[g * 32 + lane] = valid ? dg[g] : 0.0;
  // header: rows, valid lanes, uniform tails, full rows
  const uint32_t n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
  const uint32_t full_ref = __reduce_min_sync(0xFFFFFFFFu, valid ? nr / 4u : J.wr);
  const uint32_t full_alt = __reduce_min_sync(0xFFFFFFFFu, valid ? na / 4u : J.wa);
  const uint32_t nr0 = __shfl_sync(0xFFFFFFFFu, nr, 0), na0 = __shfl_sync(0xFFFFFFFFu, na, 0);
  const bool same_ref = __all_sync(0xFFFFFFFFu, !valid || nr == nr0), same_alt = __all_sync(0xFFFFFFFFu, !valid || na == na0);
  if (lane == 0) {
    const uint32_t tail_ref = (same_ref && n_valid && J.wr == full_ref + 1) ? nr0 % 4u : 0u;
    const uint32_t tail_alt = (same_alt && n_valid && J.wa == full_alt + 1) ? na0 % 4u : 0u;
    uint32_t *hdr = reinterpret_cast<uint32_t *>(blob);
    hdr[0] = J.wr; hdr[1] = J.wa; hdr[2] = n_valid | (tail_ref << 8) | (tail_alt << 12); hdr[3] = full_ref | (full_alt << 16);
  }
}

template <typename T>
struct DevBuf {  // a device array that frees itself
  T *p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    n = count;
    return count ? cudaMalloc(&p, count * sizeof(T)) : cudaSuccess;
  }
  ~DevBuf() { if (p) cudaFree(p); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

}  // namespace

// The panel as the device sees it: uploaded once, shared by every sample ingested against it.
struct vb2_panel {
  int device = 0;
  uint32_t n_marker = 0, n_pc = 0, n_chrom = 0, n_bed = 0;
  DevBuf<double> ud, means;
  DevBuf<uint8_t> alt;
  DevBuf<uint64_t> row_key, bed_keys;  // per row (chrom << 32 | pos); the same, sorted and unique
  DevBuf<char> names;
  DevBuf<uint32_t> name_off;
  DevBuf<FillTables> tables;
  double log_other[kNumQual];
};

struct vb2_ingest {
  const vb2_panel *panel = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t n_bytes = 0;
  uint32_t n_lines = 0, n_info = 0;
  unsigned long long num_bases = 0;
  DevBuf<uint8_t> text;
  DevBuf<LineRec> recs;
  DevBuf<int32_t> line_depth, row_info, row_depth;
  DevBuf<uint32_t> info_line;
  std::vector<int32_t> h_row_depth;
};

extern "C" {

int vb2_panel_create(const vb2_panel_desc *d, vb2_panel **out) {
  if (!out) return fail(VB2_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (!d || d->struct_size != sizeof(vb2_panel_desc)) return fail(VB2_ERR_INVALID, "vb2_panel_desc.struct_size mismatch");
  if (d->n_pc == 0 || d->n_pc > VB2_MAX_PC || d->ud_stride < d->n_pc) return fail(VB2_ERR_INVALID, "bad n_pc / ud_stride");
  if (d->n_marker && (!d->ud || !d->means || !d->chrom_id || !d->pos || !d->alt_base || !d->chrom_names))
    return fail(VB2_ERR_INVALID, "null array in vb2_panel_desc");
  ING_CUDA(cudaSetDevice(d->device));
  vb2_panel *P = new (std::nothrow) vb2_panel();
  if (!P) return fail(VB2_ERR_NOMEM, "out of host memory");
  struct Guard { vb2_panel *p; ~Guard() { delete p; } } guard{P};
  P->device = d->device; P->n_marker = d->n_marker; P->n_pc = d->n_pc; P->n_chrom = d->n_chrom;
  // chromosome names: offsets of the NUL-terminated strings
  std::vector<uint32_t> off(d->n_chrom + 1, 0);
  for (uint32_t c = 0; c < d->n_chrom; ++c) off[c + 1] = off[c] + (uint32_t)strlen(d->chrom_names + off[c]) + 1u;
  std::vector<uint64_t> row_key(d->n_marker);
  for (uint32_t i = 0; i < d->n_marker; ++i) {
    if (d->chrom_id[i] >= d->n_chrom) return fail(VB2_ERR_INVALID, "chrom_id out of range");
    row_key[i] = ((uint64_t)d->chrom_id[i] << 32) | (uint32_t)d->pos[i];
  }
  std::vector<uint64_t> bed(row_key);
  std::sort(bed.begin(), bed.end());
  bed.erase(std::unique(bed.begin(), bed.end()), bed.end());
  P->n_bed = (uint32_t)bed.size();
  std::vector<double> ud((size_t)d->n_marker * d->n_pc);
  for (uint32_t i = 0; i < d->n_marker; ++i)
    for (uint32_t k = 0; k < d->n_pc; ++k) ud[(size_t)i * d->n_pc + k] = d->ud[(size_t)i * d->ud_stride + k];
  FillTables T;
  double phred[kNumQual];
  vb2::build_phred_table(phred);
  vb2::emission_tables(phred, T.a_ref, T.a_alt, P->log_other);
  ING_CUDA(P->ud.alloc(ud.size()));
  ING_CUDA(P->means.alloc(d->n_marker));
  ING_CUDA(P->alt.alloc(d->n_marker));
  ING_CUDA(P->row_key.alloc(d->n_marker));
  ING_CUDA(P->bed_keys.alloc(bed.size()));
  ING_CUDA(P->names.alloc(off[d->n_chrom] ? off[d->n_chrom] : 1));
  ING_CUDA(P->name_off.alloc(off.size()));
  ING_CUDA(P->tables.alloc(1));
  if (d->n_marker) {
    ING_CUDA(cudaMemcpy(P->ud.p, ud.data(), ud.size() * sizeof(double), cudaMemcpyHostToDevice));
    ING_CUDA(cudaMemcpy(P->means.p, d->means, d->n_marker * sizeof(double), cudaMemcpyHostToDevice));
    ING_CUDA(cudaMemcpy(P->alt.p, d->alt_base, d->n_marker, cudaMemcpyHostToDevice));
    ING_CUDA(cudaMemcpy(P->row_key.p, row_key.data(), row_key.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    ING_CUDA(cudaMemcpy(P->bed_keys.p, bed.data(), bed.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  }
  if (off[d->n_chrom]) ING_CUDA(cudaMemcpy(P->names.p, d->chrom_names, off[d->n_chrom], cudaMemcpyHostToDevice));
  ING_CUDA(cudaMemcpy(P->name_off.p, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  ING_CUDA(cudaMemcpy(P->tables.p, &T, sizeof(T), cudaMemcpyHostToDevice));
  guard.p = nullptr;
  *out = P;
  return VB2_OK;
}

void vb2_panel_destroy(vb2_panel *P) {
  if (!P) return;
  cudaSetDevice(P->device);
  delete P;
}

void vb2_ingest_destroy(vb2_ingest *I) {
  if (!I) return;
  cudaSetDevice(I->panel->device);
  if (I->stream) {
    cudaStreamSynchronize(I->stream);
    cudaStreamDestroy(I->stream);
  }
  delete I;
}

int vb2_ingest_parse(const vb2_panel *P, const char *text, uint64_t n_bytes, vb2_ingest **out, vb2_ingest_info *info) {
  if (!out) return fail(VB2_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (!P || (!text && n_bytes) || !info || info->struct_size != sizeof(vb2_ingest_info)) return fail(VB2_ERR_INVALID, "bad argument");
  if (n_bytes >= 0x7FFFFFF0ull) return fail(VB2_ERR_UNSUPPORTED, "pileup text of 2 GiB or more: host reader");
  ING_CUDA(cudaSetDevice(P->device));
  vb2_ingest *I = new (std::nothrow) vb2_ingest();
  if (!I) return fail(VB2_ERR_NOMEM, "out of host memory");
  struct Guard { vb2_ingest *p; ~Guard() { if (p) vb2_ingest_destroy(p); } } guard{I};
  I->panel = P;
  I->n_bytes = n_bytes;
  ING_CUDA(cudaStreamCreateWithFlags(&I->stream, cudaStreamNonBlocking));
  cudaStream_t st = I->stream;
  ING_CUDA(I->text.alloc(n_bytes + 16));
  ING_CUDA(cudaMemcpyAsync(I->text.p, text, n_bytes, cudaMemcpyHostToDevice, st));
  // ---- lines: the positions of the newlines (a last line without one still counts, as std::getline has it) ----
  DevBuf<uint32_t> newline, count;
  ING_CUDA(newline.alloc(n_bytes + 16));
  ING_CUDA(count.alloc(4));
  ING_CUDA(cudaMemsetAsync(count.p, 0, 4 * sizeof(uint32_t), st));
  size_t tmp_bytes = 0;
  thrust::counting_iterator<uint32_t> idx(0);
  IsNewline is_nl{I->text.p};
  ING_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, newline.p, count.p, (int)n_bytes, is_nl, st));
  DevBuf<uint8_t> tmp;
  ING_CUDA(tmp.alloc(tmp_bytes + 16));
  ING_CUDA(cub::DeviceSelect::If(tmp.p, tmp_bytes, idx, newline.p, count.p, (int)n_bytes, is_nl, st));
  uint32_t n_newline = 0;
  ING_CUDA(cudaMemcpyAsync(&n_newline, count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaStreamSynchronize(st));
  const bool open_tail = n_bytes > 0 && text[n_bytes - 1] != '\n';
  const uint32_t n_lines = n_newline + (open_tail ? 1u : 0u);
  I->n_lines = n_lines;
  // ---- fields, kept bases, membership in the panel ----
  DevBuf<uint32_t> matched, rank, anomaly;
  DevBuf<unsigned long long> num_bases;
  ING_CUDA(I->recs.alloc(n_lines + 1));
  ING_CUDA(I->line_depth.alloc(n_lines + 1));
  ING_CUDA(matched.alloc(n_lines + 1));
  ING_CUDA(rank.alloc(n_lines + 1));
  ING_CUDA(anomaly.alloc(1));
  ING_CUDA(num_bases.alloc(1));
  ING_CUDA(cudaMemsetAsync(anomaly.p, 0, sizeof(uint32_t), st));
  ING_CUDA(cudaMemsetAsync(num_bases.p, 0, sizeof(unsigned long long), st));
  ING_CUDA(cudaMemsetAsync(matched.p, 0, (n_lines + 1) * sizeof(uint32_t), st));
  if (n_lines)
    parse_lines_kernel<<<(n_lines + 127) / 128, 128, 0, st>>>(I->text.p, n_bytes, newline.p, n_newline, n_lines, P->names.p,
                                                              P->name_off.p, P->n_chrom, P->bed_keys.p, P->n_bed, I->recs.p,
                                                              I->line_depth.p, matched.p, anomaly.p, num_bases.p);
  // ---- info index = rank among the matched lines, in file order (cpp:800-803: globalIndex) ----
  size_t scan_bytes = 0;
  ING_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, matched.p, rank.p, (int)(n_lines + 1), st));
  DevBuf<uint8_t> tmp2;
  ING_CUDA(tmp2.alloc(scan_bytes + 16));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(tmp2.p, scan_bytes, matched.p, rank.p, (int)(n_lines + 1), st));
  uint32_t n_info = 0, h_anom = 0;
  ING_CUDA(cudaMemcpyAsync(&n_info, rank.p + n_lines, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaMemcpyAsync(&h_anom, anomaly.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaMemcpyAsync(&I->num_bases, num_bases.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaStreamSynchronize(st));
  I->n_info = n_info;
  // ---- join: sorted (key, info) pairs, duplicates are the host reader's business, then one search per panel row ----
  DevBuf<uint64_t> info_key, sorted_key;
  DevBuf<uint32_t> info_id, sorted_id;
  ING_CUDA(I->info_line.alloc(n_info + 1));
  ING_CUDA(info_key.alloc(n_info + 1));
  ING_CUDA(sorted_key.alloc(n_info + 1));
  ING_CUDA(info_id.alloc(n_info + 1));
  ING_CUDA(sorted_id.alloc(n_info + 1));
  ING_CUDA(I->row_info.alloc(P->n_marker + 1));
  ING_CUDA(I->row_depth.alloc(P->n_marker + 1));
  if (n_lines)
    scatter_info_kernel<<<(n_lines + 255) / 256, 256, 0, st>>>(matched.p, rank.p, I->recs.p, n_lines, I->info_line.p, info_key.p, info_id.p);
  if (n_info) {
    size_t sort_bytes = 0;
    ING_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, info_key.p, sorted_key.p, info_id.p, sorted_id.p, (int)n_info, 0, 64, st));
    DevBuf<uint8_t> tmp3;
    ING_CUDA(tmp3.alloc(sort_bytes + 16));
    ING_CUDA(cub::DeviceRadixSort::SortPairs(tmp3.p, sort_bytes, info_key.p, sorted_key.p, info_id.p, sorted_id.p, (int)n_info, 0, 64, st));
    dup_check_kernel<<<(n_info + 255) / 256, 256, 0, st>>>(sorted_key.p, n_info, anomaly.p);
    ING_CUDA(cudaStreamSynchronize(st));  // (tmp3 is freed at the end of this block)
  }
  if (P->n_marker)
    join_rows_kernel<<<(P->n_marker + 255) / 256, 256, 0, st>>>(P->row_key.p, P->n_marker, sorted_key.p, sorted_id.p, n_info,
                                                                I->info_line.p, I->line_depth.p, I->row_info.p, I->row_depth.p);
  I->h_row_depth.assign(P->n_marker, -1);
  if (P->n_marker)
    ING_CUDA(cudaMemcpyAsync(I->h_row_depth.data(), I->row_depth.p, P->n_marker * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaMemcpyAsync(&h_anom, anomaly.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaStreamSynchronize(st));
  ING_CUDA(cudaGetLastError());
  if (h_anom) {
    char why[400];
    snprintf(why, sizeof(why), "pileup text needs the host reader (anomaly mask 0x%x: 1 short line, 2 non-numeric field, 4 '.' "
             "reference, 8 missing qualities, 16 indel length, 32 duplicated position, 64 > 65535 reads)", h_anom);
    return fail(VB2_ERR_UNSUPPORTED, why);
  }
  info->n_lines = n_lines;
  info->n_matched = n_info;
  info->num_bases = I->num_bases;
  info->row_depth = I->h_row_depth.data();
  guard.p = nullptr;
  *out = I;
  return VB2_OK;
}

int vb2_ingest_flatten(vb2_ingest *I, const vb2_flatten_desc *fd, vb2_llk_ctx **out) {
  if (!out) return fail(VB2_ERR_INVALID, "null out pointer");
  *out = nullptr;
  if (!I || !fd || fd->struct_size != sizeof(vb2_flatten_desc)) return fail(VB2_ERR_INVALID, "bad argument");
  const vb2_panel *P = I->panel;
  if (fd->device != P->device) return fail(VB2_ERR_INVALID, "the panel lives on another device");
  const uint32_t shard_count = fd->shard_count ? fd->shard_count : 1;
  if (fd->shard_rank >= shard_count) return fail(VB2_ERR_INVALID, "shard_rank >= shard_count");
  // (the read totals below are counted over the whole sample; shards of one sample on several GPUs go through vb2_llk_create)
  if (shard_count != 1) return fail(VB2_ERR_UNSUPPORTED, "the device ingest takes whole samples (shard_count == 1)");
  vb2_llk_ctx *ctx = vb2::ctx_new();
  if (!ctx) return fail(VB2_ERR_NOMEM, "out of host memory");
  struct Guard { vb2_llk_ctx *p; ~Guard() { if (p) vb2_llk_destroy(p); } } guard{ctx};
  const vb2::CreateParams cp{fd->device, fd->stream, fd->flags, fd->panel_dtype, fd->min_af, fd->max_af};
  int rc = vb2::ctx_open(cp, ctx);
  if (rc) { g_ingest_error = vb2_last_error(ctx); return rc; }
  const vb2::PackConfig cfg = vb2::ctx_pack_config(ctx, cp);
  cudaStream_t st = I->stream;
  const uint32_t n_rows = P->n_marker;
  // ---- 1. skip rules + class counts ----
  DevBuf<uint32_t> used, n_ref, n_alt, rank;
  DevBuf<unsigned long long> other_hist, totals;
  ING_CUDA(used.alloc(n_rows + 1));
  ING_CUDA(n_ref.alloc(n_rows + 1));
  ING_CUDA(n_alt.alloc(n_rows + 1));
  ING_CUDA(rank.alloc(n_rows + 1));
  ING_CUDA(other_hist.alloc(kNumQual));
  ING_CUDA(totals.alloc(4));
  ING_CUDA(cudaMemsetAsync(other_hist.p, 0, kNumQual * sizeof(unsigned long long), st));
  ING_CUDA(cudaMemsetAsync(totals.p, 0, 4 * sizeof(unsigned long long), st));
  ING_CUDA(cudaMemsetAsync(used.p, 0, (n_rows + 1) * sizeof(uint32_t), st));
  const double lo = fd->avg_depth - 3 * fd->sd_depth, hi = fd->avg_depth + 3 * fd->sd_depth;  // h:243-246, as the host forms them
  if (n_rows)
    marker_counts_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(I->text.p, I->recs.p, I->info_line.p, I->row_info.p, I->row_depth.p,
                                                                P->alt.p, n_rows, fd->sanity_disabled, lo, hi, used.p, n_ref.p, n_alt.p,
                                                                other_hist.p, totals.p);
  // ---- 2. used markers in panel order, their keys, the stable sort ----
  size_t scan_bytes = 0;
  ING_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, used.p, rank.p, (int)(n_rows + 1), st));
  DevBuf<uint8_t> tmp;
  ING_CUDA(tmp.alloc(scan_bytes + 16));
  ING_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, used.p, rank.p, (int)(n_rows + 1), st));
  uint32_t n_used = 0;
  unsigned long long h_hist[kNumQual], h_tot[4];
  ING_CUDA(cudaMemcpyAsync(&n_used, rank.p + n_rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaMemcpyAsync(h_hist, other_hist.p, sizeof(h_hist), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaMemcpyAsync(h_tot, totals.p, sizeof(h_tot), cudaMemcpyDeviceToHost, st));
  ING_CUDA(cudaStreamSynchronize(st));
  DevBuf<uint32_t> used_row, id, order;
  DevBuf<uint64_t> key, key_sorted;
  ING_CUDA(used_row.alloc(n_used + 1));
  ING_CUDA(id.alloc(n_used + 1));
  ING_CUDA(order.alloc(n_used + 1));
  ING_CUDA(key.alloc(n_used + 1));
  ING_CUDA(key_sorted.alloc(n_used + 1));
  DevBuf<uint8_t> tmp_sort;
  if (n_used) {
    marker_keys_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(used.p, rank.p, n_ref.p, n_alt.p, n_rows, used_row.p, key.p, id.p);
    size_t sort_bytes = 0;
    ING_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, key.p, key_sorted.p, id.p, order.p, (int)n_used, 0, 48, st));
    ING_CUDA(tmp_sort.alloc(sort_bytes + 16));
    ING_CUDA(cub::DeviceRadixSort::SortPairs(tmp_sort.p, sort_bytes, key.p, key_sorted.p, id.p, order.p, (int)n_used, 0, 48, st));
  }
  // ---- 3a. slice geometry on the device; 3b-5. cost order, shard, deal, round table on the host ----
  const uint32_t total_slices = (n_used + 31u) / 32u;
  DevBuf<vb2::SliceGeom> d_geom;
  ING_CUDA(d_geom.alloc(total_slices + 1));
  std::vector<vb2::SliceGeom> all_geom(total_slices);
  if (total_slices) {
    slice_geom_kernel<<<(total_slices + 127) / 128, 128, 0, st>>>(order.p, used_row.p, n_ref.p, n_alt.p, n_used, total_slices, d_geom.p);
    ING_CUDA(cudaMemcpyAsync(all_geom.data(), d_geom.p, total_slices * sizeof(vb2::SliceGeom), cudaMemcpyDeviceToHost, st));
    ING_CUDA(cudaStreamSynchronize(st));
  }
  vb2::PackedSample M;
  M.n_pc = P->n_pc;
  M.known_af = false;
  std::vector<vb2::SliceGeom> geom;
  std::vector<uint32_t> blob_slice;
  std::string lerr;
  if (vb2::plan_layout(all_geom, P->n_pc, false, fd->shard_rank, shard_count, cfg, &M, &geom, &blob_slice, &lerr) != VB2_OK)
    return fail(VB2_ERR_INVALID, lerr);
  // ---- 6. fill: one warp per blob ----
  std::vector<BlobJob> jobs(M.n_slices);
  for (uint32_t q = 0; q < M.n_slices; ++q) {
    const vb2::Round &R = M.rounds[q / M.n_bins];
    const vb2::SliceGeom &G = geom[blob_slice[q]];
    jobs[q] = {R.base + (uint64_t)(q % M.n_bins) * R.stride, G.first, G.wr, G.wa, 0u};
  }
  uint8_t *d_blob = nullptr;
  if (M.blob_bytes) {
    ING_CUDA(cudaMalloc(&d_blob, M.blob_bytes));
    cudaError_t e = cudaMemsetAsync(d_blob, 0xFF, M.blob_bytes, st);  // 0xFF = pad byte everywhere a read is not written
    DevBuf<BlobJob> d_jobs;
    if (e == cudaSuccess) e = d_jobs.alloc(jobs.size());
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(BlobJob), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
      const dim3 grid((M.n_slices + kFillWarps - 1) / kFillWarps);
      if (cfg.panel_fp64)
        fill_blobs_kernel<true><<<grid, 32 * kFillWarps, 0, st>>>(I->text.p, I->recs.p, I->info_line.p, I->row_info.p, P->alt.p, P->ud.p,
                                                                  P->n_pc, P->means.p, order.p, used_row.p, n_used, d_jobs.p, M.n_slices,
                                                                  P->n_pc, M.layout, P->tables.p, d_blob);
      else
        fill_blobs_kernel<false><<<grid, 32 * kFillWarps, 0, st>>>(I->text.p, I->recs.p, I->info_line.p, I->row_info.p, P->alt.p, P->ud.p,
                                                                   P->n_pc, P->means.p, order.p, used_row.p, n_used, d_jobs.p, M.n_slices,
                                                                   P->n_pc, M.layout, P->tables.p, d_blob);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cudaFree(d_blob);
      return fail(VB2_ERR_CUDA, std::string("fill_blobs_kernel: ") + cudaGetErrorString(e));
    }
  }
  uint64_t hist64[kNumQual];
  for (int q = 0; q < kNumQual; ++q) hist64[q] = h_hist[q];
  M.n_used = n_used;
  M.reads_used = h_tot[0];
  M.reads_streamed = h_tot[1];
  M.reads_folded = h_tot[2];
  M.log_other_const = vb2::other_const(hist64, P->log_other);
  rc = vb2::ctx_adopt_image(ctx, cp, M, d_blob);
  if (rc) { g_ingest_error = vb2_last_error(ctx); return rc; }
  guard.p = nullptr;
  *out = ctx;
  return VB2_OK;
}

int vb2_llk_debug_image(vb2_llk_ctx *ctx, void *dst, uint64_t n_bytes) {
  if (!ctx || (!dst && n_bytes)) return fail(VB2_ERR_INVALID, "bad argument");
  return vb2::ctx_read_image(ctx, static_cast<uint8_t *>(dst), n_bytes);
}

}  // extern "C"
