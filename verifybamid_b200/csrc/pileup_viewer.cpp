// pileup_viewer.cpp -- see pileup_viewer.h.
#include "pileup_viewer.h"

#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace vb2 {

void ParsePileupSeqBasesOnly(const std::string &seq, const std::string &qual, std::string &pseq, std::string &pqual) {
  pseq.clear();
  pqual.clear();
  auto next_qual = [&](size_t iq) -> char {
    if (iq >= qual.size()) throw std::runtime_error("Pileup format error: fewer qualities than bases");
    return qual[iq];
  };
  size_t iq = 0;
  for (size_t i = 0; i < seq.size(); ++i) {
    const char c = seq[i];
    if (c == '+' || c == '-') {
      // an indel: "+<n><n bases>" -- skipped, consumes no quality (cpp:716-723)
      size_t t = i + 1;
      while (t != seq.size() && std::isdigit((unsigned char)seq[t])) ++t;
      const size_t digitLen = t - (i + 1);
      const int clipLen = std::stoi(seq.substr(i + 1, digitLen));  // throws like the reference on "+x"
      i += digitLen + clipLen;
    } else if (c == '^') {
      i += 1;  // read start: the next char is a mapping quality (cpp:724-726)
    } else if (c == '.' || c == ',' || c == 'A' || c == 'G' || c == 'C' || c == 'T' || c == 'N' || c == 'a' ||
               c == 'g' || c == 'c' || c == 't' || c == 'n') {
      pseq += c;
      pqual += next_qual(iq);
      ++iq;
    } else if (c == '*' || c == '#') {
      ++iq;  // deletion placeholder: not modelled, but it owns a quality (cpp:738-743)
    }
    // anything else ('$', '<', '>', ...) is ignored and owns no quality
  }
}

int SimplePileupViewer::ReadPileup(const std::string &filePath, const BED &bedTable) {
  std::ifstream fin(filePath);
  numBases = 0;
  if (!fin.is_open()) throw std::runtime_error("open file " + filePath + " failed!");
  // pileup variables live across lines, exactly as in the reference: a short line leaves the
  // fields it does not supply at their previous values
  std::string pChr, refAllele, seq, qual, pileupLine, pseq, pqual;
  int pPos = 0, depth = 0;
  int globalIndex = 0;
  while (std::getline(fin, pileupLine)) {
    std::stringstream ss(pileupLine);
    ss >> pChr >> pPos >> refAllele >> depth >> seq >> qual;
    if (seq.find_first_of(".,") != std::string::npos && refAllele == ".")
      throw std::runtime_error("Pileup format error: cannot find ref allele, exit!");  // cpp:771-778
    ParsePileupSeqBasesOnly(seq, qual, pseq, pqual);
    seq = pseq;
    qual = pqual;
    depth = (int)pqual.length();  // count only SNP bases, not indels

    auto chrIt = bedTable.find(pChr);
    if (chrIt == bedTable.end()) continue;
    if (chrIt->second.find(pPos) == chrIt->second.end()) continue;

    bool existed = false;
    auto &chrIndex = posIndex[pChr];
    auto posIt = chrIndex.find(pPos);
    if (posIt != chrIndex.end()) {
      existed = true;
    } else {
      chrIndex[pPos] = globalIndex;
      globalIndex++;
    }
    if (existed) {
      // The reference warns "Merged here" but discards the merged copy (cpp:812-824): only the
      // first line's reads are kept, while numBases / effectiveNumSite still grow.
      std::cerr << "[WARNING] The pileup file has duplicated lines! Merged here" << std::endl;
    } else {
      bases.insert(bases.end(), seq.begin(), seq.end());
      quals.insert(quals.end(), qual.begin(), qual.end());
      infoOffset.push_back((int64_t)bases.size());
    }
    numBases += depth;
    depth = 0;
    seq = "";
    qual = "";
    effectiveNumSite++;
  }
  avgDepth = (double)numBases / GetNumMarker();  // cpp:831
  return 0;
}

}  // namespace vb2
