// svd_panel.h -- host side of `--RefVCF`: reference panel VCF -> genotype matrix -> (device) SVD -> .UD/.mu/.bed/.V
// Mirrors, for this path, SVDcalculator (reference SVDcalculator.{h,cpp}): ReadVcf (cpp:22-224), ProcessRefVCF
// (cpp:363-449), WriteSVD (cpp:471-513).  The decomposition itself is vb2_svd_gram (include/vb2_svd.h) on the device.
#ifndef VB2_SVD_PANEL_H_
#define VB2_SVD_PANEL_H_

#include <cstdint>
#include <string>
#include <unordered_set>
#include <vector>

namespace vb2 {

class SVDcalculator {
 public:
  // ReadVcf (cpp:22-224): plain-text or gzip-compressed VCF; PASS, bi-allelic SNPs on the included chromosomes; per sample the most
  // likely genotype from PL, else GL, else GT; markers with more than 20 % unparsed samples are dropped; a sample
  // that could not be parsed at a kept marker stays -1 in the matrix (as in the reference).
  int ReadVcf(const std::string &VcfPath, std::vector<int8_t> &genotype, int &nSamples, int &nMarkers,
              const std::unordered_set<std::string> &includeChr);
  // ProcessRefVCF (cpp:363-449).  useGramSVD is accepted for the reference's --GramSVD flag; the device path always is
  // the Gram decomposition (cpp:258-339), which the reference documents as equivalent to its JacobiSVD path.
  void ProcessRefVCF(const std::string &VcfPath, const std::unordered_set<std::string> &includeChr,
                     bool skipMinSampleCountCheck = false, int numSVDPCs = 10, bool useGramSVD = false, int device = 0);
  void WriteSVD(const std::string &Prefix, int numSVDPCs);  // cpp:471-513

  int numIndividual = 0, numMarker = 0;
  std::vector<std::vector<double>> UD, PC;  // [marker][pc], [sample][pc]  (PCtype = double, h:15)
  std::vector<std::string> Samples;
  std::vector<double> Mu;
  std::vector<std::string> chrom;           // per kept marker (BedVec / chooseBed of the reference)
  std::vector<int> pos;
  std::vector<char> refAllele, altAllele;
};

}  // namespace vb2
#endif
