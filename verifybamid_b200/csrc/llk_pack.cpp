// llk_pack.cpp -- see llk_pack.h.  Pure host C++; no CUDA here.
#include "llk_pack.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <numeric>

namespace vb2 {

void build_phred_table(double *phred94) {
  // ContaminationEstimator.h:65-74: t[i] = pow(10.0, i / -10.0)
  for (int i = 0; i < kNumQual; ++i) phred94[i] = std::pow(10.0, i / -10.0);
}

namespace {

// ContaminationEstimator.h:180-184
inline int classify_base(char base, char alt) {
  if (base == '.' || base == ',') return 0;
  if (std::toupper((unsigned char)base) == std::toupper((unsigned char)alt)) return 1;
  return 2;
}

// ContaminationEstimator.h:296-298
inline int clamp_qual(char qc) {
  int q = (int)(unsigned char)qc - 33;
  if (q < 0) q = 0;
  else if (q > 93) q = 93;
  return q;
}

struct MarkerTmp {
  uint32_t panel_row;
  int64_t beg, end;
  uint32_t n_ref, n_alt;
};

}  // namespace

int pack_sample(const vb2_llk_desc &d, const double *phred, PackedSample *out, std::string *err) {
  auto fail = [&](const char *m) { if (err) *err = m; return (int)VB2_ERR_INVALID; };
  if (d.n_pc == 0 || d.n_pc > VB2_MAX_PC) return fail("n_pc must be in [1, VB2_MAX_PC]");
  if (d.ud_stride < d.n_pc) return fail("ud_stride < n_pc");
  if (d.n_marker && (!d.ud || !d.means || !d.base_info_index || !d.alt_base || !d.info_offset))
    return fail("null array in descriptor");
  const uint32_t shard_count = d.shard_count ? d.shard_count : 1;
  if (d.shard_rank >= shard_count) return fail("shard_rank >= shard_count");

  PackedSample &P = *out;
  P = PackedSample();
  P.n_pc = d.n_pc;

  // Per-(class, q, g) single-read emission A_bc[g] = E[g]*e + N[g]*(1-e)
  // (COND_LK, ContaminationEstimator.h:164-177; same expression as h:223-224 with g1 == g2).
  double a_ref[kNumQual][3], a_alt[kNumQual][3], log_other[kNumQual];
  for (int q = 0; q < kNumQual; ++q) {
    const double e = phred[q], ne = 1.0 - e;
    a_ref[q][0] = 0.0 * e + 1.0 * ne;
    a_ref[q][1] = (1.0 / 6.0) * e + 0.5 * ne;
    a_ref[q][2] = (1.0 / 3.0) * e + 0.0 * ne;
    a_alt[q][0] = (1.0 / 3.0) * e + 0.0 * ne;
    a_alt[q][1] = (1.0 / 6.0) * e + 0.5 * ne;
    a_alt[q][2] = 0.0 * e + 1.0 * ne;
    log_other[q] = std::log((2.0 / 3.0) * e + 0.0 * ne);
  }

  // ---- 1. skip rules (h:238-249) and per-marker class counts ---------------------------------
  std::vector<MarkerTmp> used;
  used.reserve(d.n_marker);
  for (uint32_t i = 0; i < d.n_marker; ++i) {
    const int32_t idx = d.base_info_index[i];
    if (idx < 0) continue;
    const int64_t beg = d.info_offset[idx], end = d.info_offset[idx + 1];
    if (end < beg) return fail("info_offset is not non-decreasing");
    const size_t size = (size_t)(end - beg);
    if (size == 0) continue;
    if (!d.sanity_disabled &&
        (size < (d.avg_depth - 3 * d.sd_depth) || size > (d.avg_depth + 3 * d.sd_depth)))
      continue;
    if (!d.bases || !d.quals) return fail("null bases/quals with non-empty markers");
    MarkerTmp m{i, beg, end, 0, 0};
    const char alt = d.alt_base[i];
    for (int64_t j = beg; j < end; ++j) {
      const int bc = classify_base(d.bases[j], alt);
      if (bc == 0) ++m.n_ref;
      else if (bc == 1) ++m.n_alt;
    }
    used.push_back(m);
  }

  // ---- 2. order markers so that the 32 lanes of a warp run the same trip counts --------------
  auto words_of = [](uint32_t n) { return (n + kReadsPerWord - 1) / kReadsPerWord; };
  std::vector<uint32_t> order(used.size());
  std::iota(order.begin(), order.end(), 0u);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    const uint32_t wa_a = words_of(used[a].n_alt), wa_b = words_of(used[b].n_alt);
    if (wa_a != wa_b) return wa_a > wa_b;
    const uint32_t wr_a = words_of(used[a].n_ref), wr_b = words_of(used[b].n_ref);
    if (wr_a != wr_b) return wr_a > wr_b;
    return used[a].panel_row < used[b].panel_row;
  });

  // ---- 3. cut into 32-marker slices; slice s belongs to shard s % shard_count -----------------
  const size_t total_slices = (order.size() + kSliceMarkers - 1) / kSliceMarkers;
  std::vector<size_t> my_slices;
  for (size_t s = d.shard_rank; s < total_slices; s += shard_count) my_slices.push_back(s);
  P.n_slices = (uint32_t)my_slices.size();
  P.m_pad = P.n_slices * kSliceMarkers;
  P.slice_desc.assign((size_t)P.n_slices * 2, 0u);
  P.ud.assign((size_t)P.n_pc * P.m_pad, 0.0);
  P.mu.assign(P.m_pad, 1.0);
  P.diag.assign((size_t)3 * P.m_pad, 0.0);
  if (d.known_af) P.known_af.assign(P.m_pad, 0.5);
  P.marker_index.assign(P.m_pad, 0xFFFFFFFFu);

  uint64_t total_words = 0;
  for (uint32_t ls = 0; ls < P.n_slices; ++ls) {
    const size_t s = my_slices[ls];
    uint32_t wr = 0, wa = 0;
    for (size_t l = 0; l < kSliceMarkers; ++l) {
      const size_t o = s * kSliceMarkers + l;
      if (o >= order.size()) break;
      wr = std::max(wr, words_of(used[order[o]].n_ref));
      wa = std::max(wa, words_of(used[order[o]].n_alt));
    }
    if (wr > 0xFFFFu || wa > 0xFFFFu) return fail("marker deeper than 262140 reads of one class");
    if (total_words * kSliceMarkers > 0xFFFFFFFFull - (uint64_t)(wr + wa) * kSliceMarkers)
      return fail("sample too large for 32-bit word offsets (>16 GiB of reads)");
    P.slice_desc[2 * ls] = (uint32_t)(total_words * kSliceMarkers);
    P.slice_desc[2 * ls + 1] = wr | (wa << 16);
    P.max_slice_words = std::max(P.max_slice_words, wr + wa);
    total_words += wr + wa;
  }
  P.words.assign((size_t)total_words * kSliceMarkers, 0xFFFFFFFFu);
  uint8_t *bytes = reinterpret_cast<uint8_t *>(P.words.data());

  // ---- 4. fill ----------------------------------------------------------------------------------
  long double other_sum = 0.0L;
  for (uint32_t ls = 0; ls < P.n_slices; ++ls) {
    const size_t s = my_slices[ls];
    const uint32_t base = P.slice_desc[2 * ls];
    const uint32_t wr = P.slice_desc[2 * ls + 1] & 0xFFFFu;
    for (uint32_t l = 0; l < (uint32_t)kSliceMarkers; ++l) {
      const size_t o = s * kSliceMarkers + l;
      if (o >= order.size()) break;
      const MarkerTmp &m = used[order[o]];
      const uint32_t pm = ls * kSliceMarkers + l;  // packed marker id
      P.marker_index[pm] = m.panel_row;
      for (uint32_t k = 0; k < P.n_pc; ++k)
        P.ud[(size_t)k * P.m_pad + pm] = d.ud[(size_t)m.panel_row * d.ud_stride + k];
      P.mu[pm] = d.means[m.panel_row];
      if (d.known_af) P.known_af[pm] = d.known_af[m.panel_row];
      double dg[3] = {1.0, 1.0, 1.0};
      uint32_t ir = 0, ia = 0;
      const char alt = d.alt_base[m.panel_row];
      for (int64_t j = m.beg; j < m.end; ++j) {
        const int bc = classify_base(d.bases[j], alt);
        const int q = clamp_qual(d.quals[j]);
        if (bc == 2) {
          other_sum += (long double)log_other[q];
          ++P.reads_folded;
          continue;
        }
        // byte address: word t of this lane lives at words[base + t*32 + l]
        uint32_t r, t0;
        if (bc == 0) { r = ir++; t0 = 0; for (int g = 0; g < 3; ++g) dg[g] *= a_ref[q][g]; }
        else         { r = ia++; t0 = wr; for (int g = 0; g < 3; ++g) dg[g] *= a_alt[q][g]; }
        const size_t word = (size_t)base + (size_t)(t0 + r / kReadsPerWord) * kSliceMarkers + l;
        bytes[word * 4 + (r % kReadsPerWord)] = (uint8_t)q;  // little-endian: byte b = bits 8b..8b+7
        ++P.reads_streamed;
      }
      for (int g = 0; g < 3; ++g) P.diag[(size_t)g * P.m_pad + pm] = dg[g];
      P.reads_used += (uint64_t)(m.end - m.beg);
      ++P.n_used;
    }
  }
  P.log_other_const = (double)other_sum;
  return VB2_OK;
}

}  // namespace vb2

// ---------------------------------------------------------------------------------------------
// host-only diagnostics of the C ABI (include/vb2_llk.h): expose the packed image to tests
// ---------------------------------------------------------------------------------------------
extern "C" int vb2_llk_pack_host(const vb2_llk_desc *desc, vb2_packed_view *view) {
  if (!desc || !view || desc->struct_size != sizeof(vb2_llk_desc) || view->struct_size != sizeof(vb2_packed_view))
    return VB2_ERR_INVALID;
  double phred[vb2::kNumQual];
  vb2::build_phred_table(phred);
  auto *P = new vb2::PackedSample();
  std::string err;
  int rc = vb2::pack_sample(*desc, phred, P, &err);
  if (rc != VB2_OK) {
    delete P;
    return rc;
  }
  view->n_pc = P->n_pc; view->n_used = P->n_used; view->n_slices = P->n_slices; view->m_pad = P->m_pad;
  view->max_slice_words = P->max_slice_words;
  view->reads_used = P->reads_used; view->reads_streamed = P->reads_streamed; view->reads_folded = P->reads_folded;
  view->n_words = P->words.size();
  view->log_other_const = P->log_other_const;
  view->words = P->words.data(); view->slice_desc = P->slice_desc.data();
  view->ud = P->ud.data(); view->mu = P->mu.data(); view->diag = P->diag.data();
  view->known_af = P->known_af.empty() ? nullptr : P->known_af.data();
  view->marker_index = P->marker_index.data();
  view->owner = P;
  return VB2_OK;
}

extern "C" void vb2_llk_pack_free(vb2_packed_view *view) {
  if (!view || !view->owner) return;
  delete static_cast<vb2::PackedSample *>(view->owner);
  view->owner = nullptr;
}
