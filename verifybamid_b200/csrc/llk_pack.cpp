// llk_pack.cpp -- see llk_pack.h.  Pure host C++; no CUDA here.
#include "llk_pack.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

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

// Run fn(begin, end, chunk) over [0, n) on up to 8 host threads; chunk boundaries depend only on n and the
// thread count, and every caller combines per-chunk results in chunk order, so the output is deterministic.
template <typename F>
void parallel_chunks(size_t n, unsigned n_threads, F fn, size_t min_n = 4096) {
  if (n_threads <= 1 || n < min_n) {
    fn((size_t)0, n, 0u);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (n + n_threads - 1) / n_threads;
  for (unsigned t = 0; t < n_threads; ++t) {
    const size_t b = std::min(n, (size_t)t * per), e = std::min(n, b + per);
    pool.emplace_back([=, &fn]() { fn(b, e, t); });
  }
  for (auto &th : pool) th.join();
}

struct MarkerTmp {
  uint32_t panel_row;
  int64_t beg, end;
  uint32_t n_ref, n_alt;
};

}  // namespace

// Per-(class, q, g) single-read emission A_bc[g] = E[g]*e + N[g]*(1-e)  (COND_LK, ContaminationEstimator.h:164-177;
// the same expression as h:223-224 with g1 == g2), and log(2e/3) of a class-2 read.
void emission_tables(const double *phred, double (*a_ref)[3], double (*a_alt)[3], double *log_other) {
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
}
// sum over the folded (class-2) reads of log(2e/3), from their histogram over the qualities
double other_const(const uint64_t *hist, const double *log_other) {
  long double sum = 0.0L;
  for (int q = 0; q < kNumQual; ++q) sum += (long double)hist[q] * (long double)log_other[q];
  return (double)sum;
}

// Cost of a slice in quarter rows, as the kernel spends its time: a row of four real reads in every lane = 4,
// a uniform ragged tail = 2, a row that needs per-byte filler checks = 7, and 15 for the per-slice work
// (allele frequencies, priors, marginal, pipeline bookkeeping).  fr / fa = leading rows that are full in every lane.
uint32_t slice_cost(uint32_t wr, uint32_t wa, uint32_t fr, uint32_t fa, bool same_r, bool same_a) {
  const uint32_t rr = wr - fr, ra = wa - fa;
  const bool tail_r = same_r && rr == 1, tail_a = same_a && ra == 1;
  return 4u * (fr + fa) + (tail_r ? 2u : 7u * rr) + (tail_a ? 2u : 7u * ra) + 15u;
}

// Steps 3b-5 of the flatten: order the slices by cost, keep this shard's, deal them to the SM sub-partition bins in
// rounds, lay the rounds out.  Shared by the host flatten (pack_sample) and the device flatten (llk_ingest.cu), so both
// produce the same image.
int plan_layout(std::vector<SliceGeom> all_geom, uint32_t n_pc, bool known_af, uint32_t shard_rank, uint32_t shard_count,
                const PackConfig &cfg, PackedSample *out, std::vector<SliceGeom> *geom_out, std::vector<uint32_t> *blob_slice_out,
                std::string *err) {
  PackedSample &P = *out;
  // ---- 3b. costliest first; slice s of that order belongs to shard s % shard_count ----
  std::stable_sort(all_geom.begin(), all_geom.end(), [](const SliceGeom &a, const SliceGeom &b) {
    return a.cost != b.cost ? a.cost > b.cost : a.wr + a.wa > b.wr + b.wa;
  });
  std::vector<SliceGeom> &geom = *geom_out;
  geom.clear();
  for (size_t s = shard_rank; s < all_geom.size(); s += shard_count) geom.push_back(all_geom[s]);
  P.n_slices = (uint32_t)geom.size();
  // ---- 4. blob layout ------------------------------------------------------------------------------
  BlobLayout &L = P.layout;
  L = BlobLayout();
  L.panel_elem = cfg.panel_fp64 ? 8u : 4u;
  uint32_t off = kBlobHeaderBytes;
  if (known_af) {
    L.off_kaf = off; off += kSliceMarkers * 8u;
  } else {
    L.off_ud = off;  off += n_pc * kSliceMarkers * L.panel_elem;
    L.off_mu = off;  off += kSliceMarkers * L.panel_elem;
    off = (off + 7u) & ~7u;
  }
  L.off_diag = off;  off += 3u * kSliceMarkers * 8u;
  L.off_words = off;

  // ---- 5. deal the slices to SM sub-partition bins in rounds ------------------------------------------
  // Round r = the slices [r * n_bins, (r+1) * n_bins) of the cost order, one per bin.  Inside a round the
  // costliest slice goes to the bin that has the least work so far (so the bins stay level whatever the cost
  // profile is); a partial last round is dealt first, so the bins that own one blob more start with that handicap.  Bins are only labels, so they
  // are renumbered at the end to put the bins that own a blob in the last round last: in every round the bins
  // that own a blob are then a contiguous range and blob k of the round belongs to bin first_bin + k.
  P.grid_x = std::max(1u, std::min(cfg.max_ctas ? cfg.max_ctas : 1u, (P.n_slices + kBinsPerCta - 1) / kBinsPerCta));
  if (cfg.min_rounds > 1) P.grid_x = std::max(1u, std::min(P.grid_x, P.n_slices / (kBinsPerCta * cfg.min_rounds)));
  P.n_bins = P.grid_x * kBinsPerCta;
  std::vector<uint32_t> slice_bin(P.n_slices, 0);  // provisional bin label of slice j
  {
    std::vector<uint64_t> load(P.n_bins, 0);
    std::vector<uint32_t> by_load(P.n_bins);
    std::iota(by_load.begin(), by_load.end(), 0u);
    // the partial last round first: its slices are the handicap of the bins that will own six instead of five
    const uint32_t n_full_rounds = P.n_slices / P.n_bins;
    for (uint32_t j = n_full_rounds * P.n_bins, i = 0; j < P.n_slices; ++j, ++i) {
      slice_bin[j] = i;
      load[i] += geom[j].cost;
    }
    for (uint32_t r = 0; r < n_full_rounds; ++r) {
      const uint32_t j0 = r * P.n_bins;
      std::stable_sort(by_load.begin(), by_load.end(), [&](uint32_t a, uint32_t b) { return load[a] < load[b]; });
      for (uint32_t i = 0; i < P.n_bins; ++i) {  // slice j0 + i is the (i+1)-th costliest of the round
        slice_bin[j0 + i] = by_load[i];
        load[by_load[i]] += geom[j0 + i].cost;
      }
    }
    // renumber: bins without a blob in the (partial) last round first, in their old order
    const uint32_t rem = P.n_slices % P.n_bins;
    if (rem) {
      std::vector<uint8_t> in_last(P.n_bins, 0);
      for (uint32_t j = P.n_slices - rem; j < P.n_slices; ++j) in_last[slice_bin[j]] = 1;
      std::vector<uint32_t> relabel(P.n_bins);
      uint32_t next_short = 0, next_long = P.n_bins - rem;
      for (uint32_t b = 0; b < P.n_bins; ++b) relabel[b] = in_last[b] ? next_long++ : next_short++;
      for (uint32_t &b : slice_bin) b = relabel[b];
    }
  }
  // blob_slice[r * n_bins + (bin - first_bin of round r)] = the slice stored there
  std::vector<uint32_t> &blob_slice = *blob_slice_out;
  blob_slice.assign(P.n_slices, 0);
  P.rounds.clear();
  P.max_stride = 0;
  uint64_t total_bytes = 0;
  for (uint32_t j0 = 0, r = 0; j0 < P.n_slices; j0 += P.n_bins, ++r) {
    const uint32_t cnt = std::min(P.n_bins, P.n_slices - j0);
    uint32_t wmax = 0;
    for (uint32_t j = j0; j < j0 + cnt; ++j) wmax = std::max(wmax, geom[j].wr + geom[j].wa);
    const uint64_t stride64 = (uint64_t)L.off_words + (uint64_t)wmax * 128u;  // off_words % 16 == 0
    if (stride64 > 0x7FFFFFF0ull) {
      if (err) *err = "marker too deep: one blob would exceed 2 GiB";
      return VB2_ERR_INVALID;
    }
    for (uint32_t j = j0; j < j0 + cnt; ++j)
      if (geom[j].wr > 0xFFFFu || geom[j].wa > 0xFFFFu) {  // (the blob header keeps the full-row counts in 16 bits each)
        if (err) *err = "marker too deep: more than 262,140 reads of one allele on a site";
        return VB2_ERR_INVALID;
      }
    Round R{total_bytes, (uint32_t)stride64, P.n_bins - cnt, cnt, wmax};
    for (uint32_t j = j0; j < j0 + cnt; ++j) blob_slice[j0 + (slice_bin[j] - R.first_bin)] = j;
    P.rounds.push_back(R);
    P.max_stride = std::max(P.max_stride, R.stride);
    total_bytes += (uint64_t)R.stride * cnt;
    total_bytes = (total_bytes + 127u) & ~127ull;
  }
  P.conc_rounds = std::max(1u, std::min((uint32_t)P.rounds.size(), kMaxConcRounds));
  P.blob_bytes = total_bytes;
  return VB2_OK;
}

int pack_sample(const vb2_llk_desc &d, const PackConfig &cfg, const double *phred, PackedSample *out,
                std::string *err) {
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
  const bool timing = getenv("VB2_LLK_PACK_TIMING") != nullptr;
  auto tick = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "pack: %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
    tick = now;
  };
  P.known_af = d.known_af != nullptr;

  double a_ref[kNumQual][3], a_alt[kNumQual][3], log_other[kNumQual];
  emission_tables(phred, a_ref, a_alt, log_other);

  // ---- 1. skip rules (h:238-249) and per-marker class counts ---------------------------------
  unsigned n_threads = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char *t = getenv("VB2_LLK_PACK_THREADS")) n_threads = (unsigned)std::max(1, atoi(t));
  std::vector<std::vector<MarkerTmp>> used_parts(n_threads);
  std::vector<const char *> part_err(n_threads, nullptr);
  parallel_chunks(d.n_marker, n_threads, [&](size_t i0, size_t i1, unsigned t) {
    std::vector<MarkerTmp> &mine = used_parts[t];
    mine.reserve(i1 - i0);
    for (size_t i = i0; i < i1; ++i) {
      const int32_t idx = d.base_info_index[i];
      if (idx < 0) continue;
      if (d.n_info > 0 && (int64_t)idx >= d.n_info) { part_err[t] = "base_info_index beyond n_info"; return; }
      const int64_t beg = d.info_offset[idx], end = d.info_offset[idx + 1];
      if (end < beg) { part_err[t] = "info_offset is not non-decreasing"; return; }
      const size_t size = (size_t)(end - beg);
      if (size == 0) continue;
      if (!d.sanity_disabled &&
          (size < (d.avg_depth - 3 * d.sd_depth) || size > (d.avg_depth + 3 * d.sd_depth)))
        continue;
      if (!d.bases || !d.quals) { part_err[t] = "null bases/quals with non-empty markers"; return; }
      MarkerTmp m{(uint32_t)i, beg, end, 0, 0};
      const char alt = d.alt_base[i];
      for (int64_t j = beg; j < end; ++j) {
        const int bc = classify_base(d.bases[j], alt);
        if (bc == 0) ++m.n_ref;
        else if (bc == 1) ++m.n_alt;
      }
      mine.push_back(m);
    }
  });
  for (const char *e : part_err)
    if (e) return fail(e);
  std::vector<MarkerTmp> used;
  used.reserve(d.n_marker);
  for (auto &part : used_parts) used.insert(used.end(), part.begin(), part.end());
  used_parts.clear();

  lap("1 skip rules + class counts");
  // ---- 2. order markers so that the 32 lanes of a warp run the same trip counts --------------
  // A warp executes max(ref words) + max(alt words) over its lanes, so markers are grouped by alt
  // words and, inside a group, ordered by ref words -- alternating direction from group to group so
  // that the slice straddling a group boundary mixes similar ref depths.
  auto words_of = [](uint32_t n) { return (n + kReadsPerWord - 1) / kReadsPerWord; };
  // one key per marker: alt READS (descending), then ref READS (descending, or ascending in odd alt
  // groups); ties keep panel order because `used` is in panel order and the sort is stable.  Sorting by
  // exact read counts (word counts follow monotonically) makes most slices hold 32 markers with the SAME
  // (n_ref, n_alt), so the ragged last word of a run has the same number of reads in every lane.
  std::vector<uint64_t> keys(used.size());
  for (size_t u = 0; u < used.size(); ++u) {
    const uint64_t na = std::min<uint64_t>(used[u].n_alt, 0xFFFFFFull), nr = std::min<uint64_t>(used[u].n_ref, 0xFFFFFFull);
    keys[u] = ((0xFFFFFFull - na) << 24) | ((na & 1u) ? nr : 0xFFFFFFull - nr);
  }
  std::vector<uint32_t> order(used.size());
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  lap("2 sort markers");
  // ---- 3. cut into 32-marker slices ---------------------------------------------------------------
  const size_t total_slices = (order.size() + kSliceMarkers - 1) / kSliceMarkers;
  std::vector<SliceGeom> all_geom(total_slices);
  for (size_t s = 0; s < total_slices; ++s) {
    uint32_t wr = 0, wa = 0, fr = 0xFFFFFFFFu, fa = 0xFFFFFFFFu, nr0 = 0, na0 = 0;
    bool same_r = true, same_a = true;
    for (size_t l = 0; l < (size_t)kSliceMarkers; ++l) {
      const size_t o = s * kSliceMarkers + l;
      if (o >= order.size()) break;
      const MarkerTmp &m = used[order[o]];
      wr = std::max(wr, words_of(m.n_ref));
      wa = std::max(wa, words_of(m.n_alt));
      fr = std::min(fr, m.n_ref / kReadsPerWord);
      fa = std::min(fa, m.n_alt / kReadsPerWord);
      if (l == 0) { nr0 = m.n_ref; na0 = m.n_alt; }
      same_r = same_r && m.n_ref == nr0;
      same_a = same_a && m.n_alt == na0;
    }
    all_geom[s] = {wr, wa, slice_cost(wr, wa, fr, fa, same_r, same_a), (uint32_t)(s * kSliceMarkers)};
  }
  lap("3 slices");
  std::vector<SliceGeom> geom;
  std::vector<uint32_t> blob_slice;
  {
    std::string lerr;
    if (plan_layout(all_geom, P.n_pc, P.known_af, d.shard_rank, shard_count, cfg, &P, &geom, &blob_slice, &lerr) != VB2_OK)
      return fail(lerr.c_str());
  }
  P.marker_index.assign((size_t)P.n_slices * kSliceMarkers, 0xFFFFFFFFu);
  P.blob.assign((size_t)P.blob_bytes, 0xFF);  // 0xFF = pad byte everywhere a read is not written
  BlobLayout &L = P.layout;

  lap("4-5 layout, rounds, alloc");
  // ---- 6. fill ------------------------------------------------------------------------------------
  struct SliceTotals { uint32_t other_hist[kNumQual] = {}; uint64_t used = 0, streamed = 0, folded = 0; uint32_t markers = 0; };
  std::vector<SliceTotals> totals(P.n_slices);
  parallel_chunks(P.n_slices, n_threads, [&](size_t q0, size_t q1, unsigned) {
  for (uint32_t q = (uint32_t)q0; q < (uint32_t)q1; ++q) {  // q = blob position: round q / n_bins, k-th blob of it
    SliceTotals &T = totals[q];
    const Round &R = P.rounds[q / P.n_bins];
    const uint32_t j = blob_slice[q];
    uint8_t *blob = P.blob.data() + R.base + (uint64_t)(q % P.n_bins) * R.stride;
    const uint32_t wr = geom[j].wr, wa = geom[j].wa;
    uint32_t n_valid = 0, full_ref = wr, full_alt = wa;  // leading rows with four real reads in EVERY lane
    uint32_t nref0 = 0xFFFFFFFFu, nalt0 = 0xFFFFFFFFu;   // read counts if identical in every valid lane
    bool same_ref = true, same_alt = true;
    // neutral values for padding lanes
    for (uint32_t l = 0; l < (uint32_t)kSliceMarkers; ++l) {
      if (P.known_af) {
        reinterpret_cast<double *>(blob + L.off_kaf)[l] = 0.5;
      } else {
        for (uint32_t k = 0; k <= P.n_pc; ++k) {  // k == n_pc: mu
          uint8_t *p = blob + (k < P.n_pc ? L.off_ud + k * kSliceMarkers * L.panel_elem : L.off_mu) + l * L.panel_elem;
          const double v = k < P.n_pc ? 0.0 : 1.0;
          if (cfg.panel_fp64) *reinterpret_cast<double *>(p) = v;
          else *reinterpret_cast<float *>(p) = (float)v;
        }
      }
      for (int g = 0; g < 3; ++g) reinterpret_cast<double *>(blob + L.off_diag)[g * kSliceMarkers + l] = 0.0;
    }
    for (uint32_t l = 0; l < (uint32_t)kSliceMarkers; ++l) {
      const size_t o = geom[j].first + l;
      if (o >= order.size()) break;
      ++n_valid;
      const MarkerTmp &m = used[order[o]];
      full_ref = std::min(full_ref, m.n_ref / kReadsPerWord);
      full_alt = std::min(full_alt, m.n_alt / kReadsPerWord);
      if (nref0 == 0xFFFFFFFFu) { nref0 = m.n_ref; nalt0 = m.n_alt; }
      same_ref = same_ref && m.n_ref == nref0;
      same_alt = same_alt && m.n_alt == nalt0;
      P.marker_index[(size_t)q * kSliceMarkers + l] = m.panel_row;
      if (P.known_af) {
        reinterpret_cast<double *>(blob + L.off_kaf)[l] = d.known_af[m.panel_row];
      } else {
        for (uint32_t k = 0; k < P.n_pc; ++k) {
          uint8_t *p = blob + L.off_ud + (k * kSliceMarkers + l) * L.panel_elem;
          const double v = d.ud[(size_t)m.panel_row * d.ud_stride + k];
          if (cfg.panel_fp64) *reinterpret_cast<double *>(p) = v;
          else *reinterpret_cast<float *>(p) = (float)v;
        }
        uint8_t *p = blob + L.off_mu + l * L.panel_elem;
        if (cfg.panel_fp64) *reinterpret_cast<double *>(p) = d.means[m.panel_row];
        else *reinterpret_cast<float *>(p) = (float)d.means[m.panel_row];
      }
      double dg[3] = {1.0, 1.0, 1.0};
      const char alt = d.alt_base[m.panel_row];
      uint8_t *wbytes = blob + L.off_words;
      // The reads of a run are stored in ascending quality (a product does not care about the order of its factors):
      // the lanes of a warp then look up nearly the same Phred errors at the same time, which turns the kernel's
      // table look-ups -- four shared-memory loads per row, the busiest pipe after FP64 -- from 3-way bank conflicts
      // into broadcasts.  Counting sort over the 94 qualities; the diagonal products follow the same order.
      uint32_t hist[2][kNumQual];
      std::memset(hist, 0, sizeof(hist));
      for (int64_t jj = m.beg; jj < m.end; ++jj) {
        const int bc = classify_base(d.bases[jj], alt);
        const int q = clamp_qual(d.quals[jj]);
        if (bc == 2) {
          ++T.other_hist[q];
          ++T.folded;
        } else {
          ++hist[bc][q];
        }
      }
      for (int bc = 0; bc < 2; ++bc) {
        const uint32_t t0 = bc == 0 ? 0u : wr;
        uint32_t r = 0;
        for (int q = 0; q < kNumQual; ++q)
          for (uint32_t c = hist[bc][q]; c; --c, ++r) {
            for (int g = 0; g < 3; ++g) dg[g] *= bc == 0 ? a_ref[q][g] : a_alt[q][g];
            // row t = t0 + r/4 of lane l; little-endian: byte b of the word = bits 8b..8b+7
            wbytes[((size_t)(t0 + r / kReadsPerWord) * kSliceMarkers + l) * 4 + (r % kReadsPerWord)] = (uint8_t)q;
            ++T.streamed;
          }
      }
      for (int g = 0; g < 3; ++g) reinterpret_cast<double *>(blob + L.off_diag)[g * kSliceMarkers + l] = dg[g];
      T.used += (uint64_t)(m.end - m.beg);
      ++T.markers;
    }
    // uniform ragged tail: every valid lane holds exactly `tail` (1..3) reads in the one row after the full rows
    const uint32_t tail_ref = (same_ref && n_valid && wr == full_ref + 1) ? nref0 % kReadsPerWord : 0u;
    const uint32_t tail_alt = (same_alt && n_valid && wa == full_alt + 1) ? nalt0 % kReadsPerWord : 0u;
    uint32_t hdr[4] = {wr, wa, n_valid | (tail_ref << 8) | (tail_alt << 12), full_ref | (full_alt << 16)};
    std::memcpy(blob, hdr, sizeof(hdr));
  }
  }, 64);
  uint64_t other_hist[kNumQual] = {};
  for (const SliceTotals &T : totals) {
    for (int q = 0; q < kNumQual; ++q) other_hist[q] += T.other_hist[q];
    P.reads_used += T.used; P.reads_streamed += T.streamed; P.reads_folded += T.folded; P.n_used += T.markers;
  }
  P.log_other_const = other_const(other_hist, log_other);
  lap("6 fill blobs");
  return VB2_OK;
}

}  // namespace vb2

// ---------------------------------------------------------------------------------------------
// host-only diagnostics of the C ABI (include/vb2_llk.h): expose the packed image to tests
// ---------------------------------------------------------------------------------------------
extern "C" int vb2_llk_pack_host(const vb2_llk_desc *desc, uint32_t max_ctas, vb2_packed_view *view) {
  if (!desc || !view || desc->struct_size != sizeof(vb2_llk_desc) || view->struct_size != sizeof(vb2_packed_view))
    return VB2_ERR_INVALID;
  double phred[vb2::kNumQual];
  vb2::build_phred_table(phred);
  auto *P = new vb2::PackedSample();
  std::string err;
  vb2::PackConfig cfg;
  cfg.max_ctas = max_ctas ? max_ctas : 148u;
  cfg.panel_fp64 = desc->panel_dtype == VB2_PANEL_FP64;
  if (desc->flags & VB2_FLAG_BATCHED) cfg.min_rounds = 5;
  int rc = vb2::pack_sample(*desc, cfg, phred, P, &err);
  if (rc != VB2_OK) {
    delete P;
    return rc;
  }
  static_assert(sizeof(vb2::Round) == 24, "vb2_packed_view.rounds layout");
  view->n_pc = P->n_pc; view->n_used = P->n_used; view->n_slices = P->n_slices;
  view->grid_x = P->grid_x; view->n_bins = P->n_bins; view->conc_rounds = P->conc_rounds;
  view->n_rounds = (uint32_t)P->rounds.size();
  view->max_stride = P->max_stride; view->known_af = P->known_af ? 1u : 0u;
  view->panel_elem = P->layout.panel_elem;
  view->off_ud = P->layout.off_ud; view->off_mu = P->layout.off_mu; view->off_kaf = P->layout.off_kaf;
  view->off_diag = P->layout.off_diag; view->off_words = P->layout.off_words;
  view->reads_used = P->reads_used; view->reads_streamed = P->reads_streamed; view->reads_folded = P->reads_folded;
  view->blob_bytes = P->blob.size();
  view->log_other_const = P->log_other_const;
  view->blob = P->blob.data();
  view->rounds = reinterpret_cast<const uint32_t *>(P->rounds.data());
  view->marker_index = P->marker_index.data();
  view->owner = P;
  return VB2_OK;
}

extern "C" void vb2_llk_pack_free(vb2_packed_view *view) {
  if (!view || !view->owner) return;
  delete static_cast<vb2::PackedSample *>(view->owner);
  view->owner = nullptr;
}
