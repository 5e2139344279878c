// llk_pack.h -- host-side flatten of the reference's per-marker base/qual vectors into the
// image the sm_100a likelihood kernel streams from HBM.
//
// What is folded in here, once per sample, because it does not change between evaluations
// (reference file:line in parentheses):
//   * the marker skip rules                     (ContaminationEstimator.h:238-249)
//   * classifyBase + the [0,93] quality clamp   (ContaminationEstimator.h:180-184, :296-298)
//   * class-2 ("other") reads: every one multiplies all nine genotype pairs by the same
//     2e/3 (COND_LK[1][g][2] = 2/3 for every g, h:173-175), so they leave the per-evaluation
//     stream and become one scalar, log_other_const
//   * the three diagonal genotype pairs g1 == g2: alpha*A[g] + (1-alpha)*A[g] = A[g] does not
//     depend on alpha or the PCs -> three per-marker constants diag[g]
//
// Layout ("blobs in rounds", DESIGN.md "Data layout in HBM"):
//   markers are sorted by (alt reads, ref reads) so the 32 lanes of a warp run equal trip
//   counts, cut into 32-marker SLICES, and the slices are dealt to BINS -- one bin per SM
//   sub-partition of the launch (4 per CTA, one CTA per SM) -- in ROUNDS, costliest first, the
//   costliest slice of a round to the least loaded bin, so every sub-partition gets the same
//   amount of work (cost = what the kernel spends: rows, ragged rows, per-slice set-up).  Everything a
//   warp needs for one slice -- header, UD columns, mu, diag, the packed quality bytes -- is ONE
//   contiguous BLOB, and all blobs of a round have the same stride, so a warp finds its blob
//   with arithmetic only (no descriptor load) and fetches it with one TMA bulk copy.
#ifndef VB2_LLK_PACK_H_
#define VB2_LLK_PACK_H_

#include <cstdint>
#include <string>
#include <vector>

#include "vb2_llk.h"

namespace vb2 {

constexpr int kSliceMarkers = 32;       // one warp = one slice = 32 markers, one per lane
constexpr int kReadsPerWord = 4;        // four quality bytes per 32-bit word
constexpr uint8_t kPadByte = 0xFF;      // "no read" filler inside a lane's last word
constexpr int kNumQual = 94;            // Phred 0..93 (h:60-74)
constexpr uint32_t kBlobHeaderBytes = 16;

constexpr uint32_t kBinsPerCta = 4;     // SM sub-partitions: warp w of a CTA issues on SMSP w % 4
constexpr uint32_t kMaxConcRounds = 4;  // rounds of a bin a one-evaluation CTA keeps in flight (4*4 = 16 warps)

struct PackConfig {
  uint32_t max_ctas = 148;   // CTAs of one launch = SMs of the device (one persistent CTA per SM)
  bool panel_fp64 = false;   // UD / mu element type inside the blobs
  uint32_t min_rounds = 1;   // use fewer CTAs' worth of bins if needed to give every bin about this many slices
};

// Byte offsets inside a blob (identical for every blob of a sample).
struct BlobLayout {
  uint32_t panel_elem = 4;  // sizeof(float) or sizeof(double)
  uint32_t off_ud = 0;      // [n_pc][32] panel_elem   (absent when known_af)
  uint32_t off_mu = 0;      // [32] panel_elem         (absent when known_af)
  uint32_t off_kaf = 0;     // [32] double             (only when known_af)
  uint32_t off_diag = 0;    // [3][32] double
  uint32_t off_words = 0;   // [wr + wa][32] uint32: row t, lane l at off_words + (t*32 + l)*4
};

struct Round {
  uint64_t base;       // byte offset of the round's first blob
  uint32_t stride;     // bytes per blob in this round (multiple of 16)
  uint32_t first_bin;  // bins [first_bin, first_bin + count) own a blob in this round;
  uint32_t count;      //   bin b's blob is at base + (b - first_bin) * stride
  uint32_t rows;       // word rows per blob = (stride - off_words) / 128
};

struct PackedSample {
  uint32_t n_pc = 0;
  uint32_t n_used = 0;          // markers of this shard that survive the skip rules
  uint32_t n_slices = 0;        // ceil(n_used / 32) for this shard
  uint32_t grid_x = 0;          // CTAs of a launch: min(max_ctas, ceil(n_slices / 4))
  uint32_t n_bins = 0;          // 4 * grid_x
  uint32_t conc_rounds = 0;     // min(rounds, kMaxConcRounds): a CTA has 4 * conc_rounds warps
  uint32_t max_stride = 0;      // largest blob stride (bytes)
  bool known_af = false;
  uint64_t reads_used = 0, reads_streamed = 0, reads_folded = 0;
  uint64_t blob_bytes = 0;      // size of the image
  double log_other_const = 0.0;
  BlobLayout layout;
  std::vector<Round> rounds;
  // Blob header (16 bytes): u32 wr (ref word rows), u32 wa (alt word rows),
  // u32 n_valid | tail_ref << 8 | tail_alt << 12 (lanes 0..n_valid-1 hold markers; tail_x = 1..3 when the single
  // row after the full rows holds exactly that many reads in EVERY valid lane, 0 = unknown / mixed),
  // u32 full_ref | full_alt << 16 (leading ref / alt rows in which every valid lane has four real reads:
  // the kernel streams those -- and uniform tails -- without looking for filler bytes).
  std::vector<uint8_t> blob;          // the whole image
  std::vector<uint32_t> marker_index; // [n_slices*32] panel row per (blob, lane) in image order: blob q is the
                                      // (q % n_bins)-th blob of round q / n_bins, i.e. of bin first_bin + q % n_bins
};

// Geometry of one 32-marker slice: ref / alt word rows, its cost, and where its markers start in the sorted marker order.
struct SliceGeom {
  uint32_t wr, wa, cost, first;
};
uint32_t slice_cost(uint32_t wr, uint32_t wa, uint32_t full_ref, uint32_t full_alt, bool same_ref, bool same_alt);
// Cost order, sharding, the deal to bins, the round table (everything of the layout that is not per-read work); shared
// by the host flatten and the device flatten.  all_geom[s] = slice s of the sorted marker order.
int plan_layout(std::vector<SliceGeom> all_geom, uint32_t n_pc, bool known_af, uint32_t shard_rank, uint32_t shard_count,
                const PackConfig &cfg, PackedSample *out, std::vector<SliceGeom> *geom, std::vector<uint32_t> *blob_slice,
                std::string *err);
void emission_tables(const double *phred, double (*a_ref)[3], double (*a_alt)[3], double *log_other);
double other_const(const uint64_t *hist, const double *log_other);

// Returns VB2_OK or VB2_ERR_INVALID (message in *err).  phred[q] = 10^(-q/10), q = 0..93.
int pack_sample(const vb2_llk_desc &d, const PackConfig &cfg, const double *phred, PackedSample *out,
                std::string *err);

// phred table exactly as the reference builds it (ContaminationEstimator.h:65-74).
void build_phred_table(double *phred94);

}  // namespace vb2
#endif
