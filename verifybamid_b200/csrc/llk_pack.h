// llk_pack.h -- host-side flatten of the reference's per-marker base/qual vectors into the
// SoA image the sm_100a likelihood kernel streams from HBM.
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
// See DESIGN.md "Data layout in HBM".
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

struct PackedSample {
  uint32_t n_pc = 0;
  uint32_t n_used = 0;          // markers of this shard that survive the skip rules
  uint32_t n_slices = 0;        // ceil(n_used / 32) for this shard
  uint32_t m_pad = 0;           // n_slices * 32
  uint32_t max_slice_words = 0; // max over slices of (ref words + alt words) per lane
  uint64_t reads_used = 0, reads_streamed = 0, reads_folded = 0;
  double log_other_const = 0.0;

  // words[slice_desc[2s] + t*32 + lane]: t-th word of lane `lane` of slice s; the first
  // wr = slice_desc[2s+1] & 0xFFFF words hold ref-class reads, the next
  // wa = slice_desc[2s+1] >> 16 hold alt-class reads; each byte is a clamped quality or 0xFF.
  std::vector<uint32_t> words;
  std::vector<uint32_t> slice_desc;   // [n_slices][2]
  std::vector<double> ud;             // [n_pc][m_pad]  (column-major: coalesced per PC)
  std::vector<double> mu;             // [m_pad]
  std::vector<double> diag;           // [3][m_pad]
  std::vector<double> known_af;       // [m_pad] or empty
  std::vector<uint32_t> marker_index; // [m_pad] panel row of each packed marker (0xFFFFFFFF pad)
};

// Returns VB2_OK or VB2_ERR_INVALID (message in *err).  phred[q] = 10^(-q/10), q = 0..93.
int pack_sample(const vb2_llk_desc &d, const double *phred, PackedSample *out, std::string *err);

// phred table exactly as the reference builds it (ContaminationEstimator.h:65-74).
void build_phred_table(double *phred94);

}  // namespace vb2
#endif
