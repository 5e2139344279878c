// svd_panel.cpp -- see svd_panel.h.  Host glue around vb2_svd_gram; reference file:line in the comments.
#include "svd_panel.h"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "estimator.h"  // notice / warning / error
#include "vb2_svd.h"

namespace vb2 {
namespace {

// statgen String::ReplaceColumns: split on ONE delimiter, empty fields kept (libVcfFile.cpp:479, :556, :573, :877, :931)
void split_columns(const std::string &s, char delim, std::vector<std::string> &out) {
  out.clear();
  size_t b = 0;
  for (;;) {
    const size_t e = s.find(delim, b);
    out.emplace_back(s, b, e == std::string::npos ? std::string::npos : e - b);
    if (e == std::string::npos) break;
    b = e + 1;
  }
}
// statgen StringArray::ReplaceTokens: split on ANY of the delimiters, empty tokens dropped (SVDcalculator.cpp:108, :122, :139)
void split_tokens(const std::string &s, const char *delims, std::vector<std::string> &out) {
  out.clear();
  size_t b = 0;
  while (b < s.size()) {
    const size_t e = s.find_first_of(delims, b);
    if (e != b) out.emplace_back(s, b, e == std::string::npos ? std::string::npos : e - b);
    if (e == std::string::npos) break;
    b = e + 1;
  }
}
int find_key(const std::vector<std::string> &keys, const char *k) {
  for (size_t i = 0; i < keys.size(); ++i)
    if (keys[i] == k) return (int)i;
  return -1;
}
void upper(std::string &s) {
  for (char &c : s) c = (char)toupper((unsigned char)c);
}

// one line at a time from a plain or gzip-compressed file (a panel VCF line carries thousands of samples)
class LineReader {
 public:
  explicit LineReader(const std::string &path) : f_(gzopen(path.c_str(), "rb")), buf_(1 << 20) {
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~LineReader() {
    if (f_) gzclose(f_);
  }
  bool ok() const { return f_ != nullptr; }
  bool getline(std::string &line) {
    line.clear();
    for (;;) {
      if (!gzgets(f_, buf_.data(), (int)buf_.size())) return !line.empty();
      const size_t n = strlen(buf_.data());
      line.append(buf_.data(), n);
      if (n && line.back() == '\n') {
        line.pop_back();
        return true;
      }
      if (gzeof(f_)) return true;
    }
  }

 private:
  gzFile f_;
  std::vector<char> buf_;
};

// libvb2svd.so sits next to the executable / libvb2llk.so; it is loaded only when --RefVCF asks for it, so the
// likelihood path never maps cuSOLVER.
typedef int (*svd_gram_fn)(const vb2_svd_desc *);
typedef const char *(*svd_err_fn)(void);
void load_svd_library(svd_gram_fn *gram, svd_err_fn *err) {
  std::string dir;
  Dl_info info;
  if (dladdr((void *)&load_svd_library, &info) && info.dli_fname) {
    dir = info.dli_fname;
    const size_t slash = dir.rfind('/');
    dir = slash == std::string::npos ? std::string(".") : dir.substr(0, slash);
  }
  const char *env = getenv("VB2_SVD_LIBRARY");
  const std::string path = env ? std::string(env) : dir + "/libvb2svd.so";
  void *h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) error("--RefVCF: cannot load %s (%s); there is no CPU fallback", path.c_str(), dlerror());
  *gram = (svd_gram_fn)dlsym(h, "vb2_svd_gram");
  *err = (svd_err_fn)dlsym(h, "vb2_svd_last_error");
  if (!*gram || !*err) error("--RefVCF: %s does not export vb2_svd_gram", path.c_str());
}

}  // namespace

int SVDcalculator::ReadVcf(const std::string &VcfPath, std::vector<int8_t> &genotype, int &nSamples, int &nMarkers,
                           const std::unordered_set<std::string> &includeChr) {
  const long maxPhred = 255;  // cpp:27
  LineReader in(VcfPath);  // plain text or gzip / bgzip (zlib reads both)
  if (!in.ok()) error("Failed to open VCF file %s", VcfPath.c_str());
  std::string line;
  std::vector<std::string> cols, alts, filters, fmt, vals, three, alleles;
  // header (libVcfFile.cpp:238-330): meta lines, then #CHROM ... FORMAT sample1 sample2 ...
  bool have_header = false;
  while (in.getline(line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.compare(0, 2, "##") == 0) continue;
    if (line.compare(0, 6, "#CHROM") == 0) {
      split_columns(line, '\t', cols);
      for (size_t i = 9; i < cols.size(); ++i) Samples.push_back(cols[i]);
      have_header = true;
      break;
    }
    error("Header line is not found : #CHROM...");
  }
  if (!have_header) error("Header line is not found : #CHROM...");
  nSamples = (int)Samples.size();
  if (nSamples == 0) error("No individual genotype information exist in the input VCF file %s", VcfPath.c_str());  // cpp:36-39
  const bool filterByChrom = !includeChr.empty();
  if (filterByChrom) notice("Filtering to %d chromosome(s) specified by --IncludeChr", (int)includeChr.size());
  nMarkers = 0;
  std::string markerName, prevMarkerName;
  std::vector<int8_t> perMarkerGeno((size_t)nSamples);
  long lineNo = 0;
  while (in.getline(line)) {
    ++lineNo;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    split_columns(line, '\t', cols);
    if (cols.size() < 8) error("VCF line with fewer than 8 columns. See line %ld.", lineNo);
    const std::string &sChrom = cols[0];
    const int nPos = atoi(cols[1].c_str());
    markerName = sChrom + ":" + std::to_string(nPos);
    if (prevMarkerName == markerName) error("Duplicated Marker: %s", markerName.c_str());  // cpp:60-63
    split_columns(cols[6], ';', filters);
    if (filters.size() > 1 || filters[0] != "PASS") {  // cpp:64-68
      warning("Skip filtered (%s) marker: %s", filters[0].c_str(), markerName.c_str());
      continue;
    }
    std::string sRef = cols[3];
    upper(sRef);
    split_columns(cols[4], ',', alts);
    for (auto &a : alts) upper(a);
    if (alts.size() > 1) {  // cpp:69-73
      warning("Skip non-Biallelic marker: %s", markerName.c_str());
      continue;
    }
    if (sRef.size() > 1 || alts[0].size() > 1) {  // cpp:74-78
      warning("Skip non-SNP marker: %s", markerName.c_str());
      continue;
    }
    if (filterByChrom && includeChr.find(sChrom) == includeChr.end()) continue;  // cpp:79-82
    if (cols.size() < 10) error("No individual genotype information exist in the input VCF file %s", VcfPath.c_str());
    split_columns(cols[8], ':', fmt);
    const int idxPL = find_key(fmt, "PL"), idxGL = find_key(fmt, "GL"), idxGT = find_key(fmt, "GT");  // cpp:90-95
    if (idxPL < 0 && idxGL < 0 && idxGT < 0) error("Cannot recognize GT, GL or PL key in FORMAT field");
    if ((int)cols.size() - 9 != nSamples) error("VCF line with %d samples, header has %d. See line %ld.", (int)cols.size() - 9, nSamples, lineNo);
    int nMissingGenoSamples = 0;
    std::fill(perMarkerGeno.begin(), perMarkerGeno.end(), (int8_t)-1);
    for (int i = 0; i < nSamples; ++i) {
      const std::string &sv = cols[9 + i];
      // libVcfFile.cpp:913-940: "./." or "." -> first value "./.", the others empty; else one value per FORMAT key
      if (sv == "./." || sv == ".") {
        vals.assign(fmt.size(), std::string());
        vals[0] = "./.";
      } else {
        split_columns(sv, ':', vals);
        if (vals.size() != fmt.size())
          error("# values = %s do not match with # fields in FORMAT field = %d at sampleIndex = %d See line %ld.", sv.c_str(),
                (int)fmt.size(), i, lineNo);
      }
      long phred11 = 0, phred12 = 0, phred22 = 0;
      bool parsed = false;
      if (!parsed && idxPL >= 0) {  // cpp:106-115
        split_tokens(vals[idxPL], ",", three);
        if (three.size() == 3 && three[0] != "." && three[1] != "." && three[2] != ".") {
          phred11 = atoi(three[0].c_str()); phred12 = atoi(three[1].c_str()); phred22 = atoi(three[2].c_str());
          parsed = true;
        }
      }
      if (!parsed && idxGL >= 0) {  // cpp:118-131
        split_tokens(vals[idxGL], ",", three);
        if (three.size() == 3 && three[0] != "." && three[1] != "." && three[2] != ".") {
          phred11 = static_cast<int>(-10. * atof(three[0].c_str()));
          phred12 = static_cast<int>(-10. * atof(three[1].c_str()));
          phred22 = static_cast<int>(-10. * atof(three[2].c_str()));
          parsed = true;
        }
      }
      if (!parsed && idxGT >= 0) {  // cpp:135-157
        split_tokens(vals[idxGT], "|/", alleles);
        if (alleles.size() == 2 && alleles[0] != "." && alleles[1] != ".") {
          const long geno = atoi(alleles[0].c_str()) + atoi(alleles[1].c_str());
          if (geno == 0) { phred11 = 0; phred12 = 30; phred22 = 50; }
          else if (geno == 1) { phred11 = 50; phred12 = 0; phred22 = 50; }
          else { phred11 = 50; phred12 = 30; phred22 = 0; }
          parsed = true;
        }
      }
      if (!parsed) {  // cpp:159-162
        nMissingGenoSamples++;
        continue;
      }
      if (phred11 < 0 || phred12 < 0 || phred22 < 0) error("Negative PL or Positive GL observed");  // cpp:164-166
      phred11 = std::min(phred11, maxPhred); phred12 = std::min(phred12, maxPhred); phred22 = std::min(phred22, maxPhred);
      int minGeno = -1;  // cpp:172-190: strictly smaller than the running minimum, which starts at maxPhred
      long minPhred = maxPhred;
      if (phred11 < minPhred) { minPhred = phred11; minGeno = 0; }
      if (phred12 < minPhred) { minPhred = phred12; minGeno = 1; }
      if (phred22 < minPhred) { minPhred = phred22; minGeno = 2; }
      perMarkerGeno[(size_t)i] = (int8_t)minGeno;
    }
    const float genoMissingRate = static_cast<float>(nMissingGenoSamples) / nSamples;  // cpp:193-198
    if (genoMissingRate > 0.2f) {
      warning("Skip marker (%s) with high missing rate (%f > 0.2) in genotype fields.", markerName.c_str(), genoMissingRate);
      continue;
    }
    genotype.insert(genotype.end(), perMarkerGeno.begin(), perMarkerGeno.end());
    chrom.push_back(sChrom);
    pos.push_back(nPos);
    refAllele.push_back(sRef[0]);
    altAllele.push_back(alts[0][0]);
    nMarkers++;
    prevMarkerName = markerName;
  }
  std::map<std::string, int> chrCounts;  // cpp:206-214
  for (const auto &c : chrom) chrCounts[c]++;
  notice("Markers retained across %d chromosome(s):", (int)chrCounts.size());
  for (const auto &kv : chrCounts) notice("  %s: %d markers", kv.first.c_str(), kv.second);
  return 0;
}

void SVDcalculator::ProcessRefVCF(const std::string &VcfPath, const std::unordered_set<std::string> &includeChr,
                                  bool skipMinSampleCountCheck, int numSVDPCs, bool useGramSVD, int device) {
  (void)useGramSVD;
  std::vector<int8_t> genotype;  // markers x samples
  ReadVcf(VcfPath, genotype, numIndividual, numMarker, includeChr);
  notice("Number of Markers after filtering: %d", numMarker);
  notice("Number of Individuals: %d", numIndividual);
  if (numMarker < 5000) error("Insufficient number of markers (need >= 5000, have %d)\n", numMarker);  // cpp:378-381
  if (numIndividual < 1000) {  // cpp:383-396
    if (skipMinSampleCountCheck)
      warning("Only %d individuals in reference panel (recommended minimum is 1000). Proceeding because "
              "--SkipMinSampleCountCheck is set. Contamination estimates may be unreliable if the panel does not adequately "
              "capture population structure.", numIndividual);
    else
      error("Insufficient number of individuals (need >= 1000, have %d). If your reference panel adequately captures "
            "population structure with fewer samples, rerun with --SkipMinSampleCountCheck.\n", numIndividual);
  }
  // cpp:411-416: numSVDPCs == 0 means "all available components"; the device call returns at most VB2_SVD_MAX_PC
  const int maxPCs = std::min(numMarker, numIndividual);
  int numPCs = numSVDPCs > 0 ? std::min(numSVDPCs, maxPCs) : maxPCs;
  if (numPCs > VB2_SVD_MAX_PC) {
    warning("--NumSVDPCs: this engine writes at most %d components (asked for %d)", VB2_SVD_MAX_PC, numPCs);
    numPCs = VB2_SVD_MAX_PC;
  }
  notice("Building genotype matrix (%d markers x %d individuals) and decomposing it on the GPU...", numMarker, numIndividual);
  svd_gram_fn gram = nullptr;
  svd_err_fn last_error = nullptr;
  load_svd_library(&gram, &last_error);
  std::vector<float> mu((size_t)numMarker), ud((size_t)numMarker * numPCs), pc((size_t)numIndividual * numPCs),
      singular((size_t)numIndividual);
  vb2_svd_timing timing = {};
  vb2_svd_desc d = {};
  d.struct_size = sizeof(d);
  d.device = device;
  d.n_marker = (uint32_t)numMarker;
  d.n_sample = (uint32_t)numIndividual;
  d.n_pc = (uint32_t)numPCs;
  d.genotype = genotype.data();
  d.mu = mu.data(); d.ud = ud.data(); d.pc = pc.data(); d.singular = singular.data();
  d.timing = &timing;
  if (gram(&d) != 0) error("GPU panel construction: %s", last_error());
  notice("Gram SVD on the device: centring %.2f ms, A^T*A %.2f ms, eigendecomposition %.2f ms, A*V %.2f ms", timing.center_ms,
         timing.gram_ms, timing.eigen_ms, timing.ud_ms);
  // logVarianceExplained (cpp:226-256), over the whole spectrum
  {
    double total = 0.0;
    for (float s : singular) total += (double)s * (double)s;
    if (total <= 0.0) {
      warning("Total variance is zero; skipping variance-explained logging.");
    } else {
      double cumulative = 0.0;
      const int numToLog = std::min(numIndividual, 20);
      for (int i = 0; i < numToLog; ++i) {
        const double sv = (double)singular[(size_t)i], ve = sv * sv / total;
        cumulative += ve;
        notice("  PC%d: singular_value=%.4f  variance_explained=%.4f (%.2f%%)  cumulative=%.4f (%.2f%%)", i + 1, sv, ve, ve * 100.0,
               cumulative, cumulative * 100.0);
      }
      if (numIndividual > numToLog) notice("  ... (%d more components not shown)", numIndividual - numToLog);
    }
  }
  Mu.assign(mu.begin(), mu.end());  // cpp:408, :431-444: float results widened to PCtype = double
  UD.assign((size_t)numMarker, std::vector<double>((size_t)numPCs, 0.0));
  for (int i = 0; i < numMarker; ++i)
    for (int j = 0; j < numPCs; ++j) UD[(size_t)i][(size_t)j] = ud[(size_t)i * numPCs + j];
  PC.assign((size_t)numIndividual, std::vector<double>((size_t)numPCs, 0.0));
  for (int i = 0; i < numIndividual; ++i)
    for (int j = 0; j < numPCs; ++j) PC[(size_t)i][(size_t)j] = pc[(size_t)i * numPCs + j];
  WriteSVD(VcfPath, numSVDPCs);
}

void SVDcalculator::WriteSVD(const std::string &Prefix, int numSVDPCs) {  // cpp:471-513: the same four files, the same formatting
  const int numAvailable = (numMarker > 0 && !UD.empty()) ? (int)UD[0].size() : 0;
  const int numToWrite = numSVDPCs <= 0 ? numAvailable : std::min(numSVDPCs, numAvailable);
  notice("Writing SVD output files with %d PCs (of %d available) to prefix: %s", numToWrite, numAvailable, Prefix.c_str());
  std::ofstream fMu(Prefix + ".mu"), fUD(Prefix + ".UD"), fPC(Prefix + ".V"), fBed(Prefix + ".bed");
  for (int i = 0; i < numMarker; ++i) {
    const std::string &chr = chrom[(size_t)i];
    const int end = pos[(size_t)i], beg = end - 1;
    fMu << chr + ":" + std::to_string(end) << "\t" << Mu[(size_t)i] << std::endl;
    fBed << chr << "\t" << beg << "\t" << end << "\t" << refAllele[(size_t)i] << "\t" << altAllele[(size_t)i] << std::endl;
    for (int j = 0; j < numToWrite; ++j) fUD << UD[(size_t)i][(size_t)j] << "\t";
    fUD << std::endl;
  }
  for (int k = 0; k < numIndividual; ++k) {
    fPC << Samples[(size_t)k] << "\t";
    for (int i = 0; i < numToWrite; ++i) fPC << PC[(size_t)k][(size_t)i] << "\t";
    fPC << std::endl;
  }
  notice("SVD output files written: %s.UD, %s.mu, %s.bed, %s.V", Prefix.c_str(), Prefix.c_str(), Prefix.c_str(), Prefix.c_str());
}

}  // namespace vb2
