// pileup_viewer.h -- text-pileup half of the reference's SimplePileupViewer
// (SimplePileupViewer.h:59-60, :94-106; SimplePileupViewer.cpp:711-833), kept as host C++.
// The per-marker base/qual vectors the reference holds as vector<vector<char>> are stored flat
// (one CSR pair) because that is exactly the image the C ABI takes (include/vb2_llk.h).
// The BAM/CRAM half (SimplePileupViewer.cpp:245-557) needs htslib and is out of scope.
#ifndef VB2_PILEUP_VIEWER_H_
#define VB2_PILEUP_VIEWER_H_

#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace vb2 {

// ContaminationEstimator.h:462-463: BED[chr][pos(1-based)] = (ref, alt)
typedef std::unordered_map<std::string, std::unordered_map<int, std::pair<char, char>>> BED;

class SimplePileupViewer {
 public:
  // viewer.baseInfo[b] / viewer.qualInfo[b] = bases/quals[infoOffset[b] .. infoOffset[b+1])
  std::vector<char> bases, quals;
  std::vector<int64_t> infoOffset{0};
  std::unordered_map<std::string, std::unordered_map<int32_t, int32_t>> posIndex;  // h:106

  std::string SEQ_SM = "DefaultSampleName";  // h:97
  long numBases = 0;                         // h:98
  int effectiveNumSite = 0;                  // h:99
  double avgDepth = 0, sdDepth = 0;          // h:100-101

  int GetNumMarker() const { return effectiveNumSite; }  // h:132-135
  size_t NumInfo() const { return infoOffset.size() - 1; }
  size_t DepthOf(int32_t infoIndex) const { return (size_t)(infoOffset[infoIndex + 1] - infoOffset[infoIndex]); }

  // SimplePileupViewer.cpp:748-833.  Throws std::runtime_error on malformed input.
  int ReadPileup(const std::string &filePath, const BED &bedTable);
};

// SimplePileupViewer.cpp:711-746: keep . , ACGTN acgtn (one qual each); skip +n/-n runs and the
// char after '^' without consuming a qual; '*' and '#' are dropped but consume a qual.
void ParsePileupSeqBasesOnly(const std::string &seq, const std::string &qual, std::string &pseq, std::string &pqual);

}  // namespace vb2
#endif
