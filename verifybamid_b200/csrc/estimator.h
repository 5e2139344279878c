// estimator.h -- host side of the likelihood path, mirroring the reference's ContaminationEstimator
// (ContaminationEstimator.h:40-542, ContaminationEstimator.cpp) with the same member names, flags and
// control flow.  The ONE thing that differs: FullLLKFunc::ComputeMixLLKs does not loop over markers on
// the CPU -- it calls the CUDA engine through the C ABI (include/vb2_llk.h).  There is no CPU path.
#ifndef VB2_ESTIMATOR_H_
#define VB2_ESTIMATOR_H_

#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "amoeba.h"
#include "pileup_viewer.h"
#include "vb2_llk.h"

namespace vb2 {

class CohortCoordinator;  // cohort.h: lock-step evaluation of many samples in one launch

void notice(const char *msg, ...);   // statgen/Error.cpp:70-79
void warning(const char *msg, ...);  // statgen/Error.cpp:42-53
[[noreturn]] void error(const char *msg, ...);  // statgen/Error.cpp:26-40: prints and throws

class ContaminationEstimator {
 public:
  bool isPCFixed = false, isAlphaFixed = false, isAFknown = false, isHeter = true;
  bool isPileupInput = false, isSanityCheckDisabled = false, verbose = false;
  int numPC = 2, numThread = 4, seed = 12345;
  double epsilon = 1e-8;
  // engine options (not in the reference)
  int numGPU = 1;           // marker shards, one per device 0..numGPU-1
  int firstDevice = 0;
  bool panelFp64 = false;   // keep UD/mu in fp64 in HBM
  bool quiet = false;       // cohort mode: no per-sample lines on stdout
  CohortCoordinator *cohort = nullptr;  // cohort mode: evaluations go through the coordinator
  int cohortIndex = -1;

  // ContaminationEstimator.h:76-443
  class FullLLKFunc : public VectorFunc {
   public:
    double llk1 = 0, llk0 = 0;
    ContaminationEstimator *ptr = nullptr;
    std::vector<double> fixPC, fixPC2, globalPC, globalPC2;
    double fixAlpha = 0, globalAlpha = 0;
    long evalCount = 0;

    static double InvLogit(double x);  // h:119-122
    static double Logit(double x);     // h:124-127
    // h:194-314 -- evaluated on the GPU(s): sum of the marker shards' partial log-likelihoods
    double ComputeMixLLKs(const std::vector<double> &tPC1, const std::vector<double> &tPC2, double alpha);
    int Initialize();     // h:316-332
    int CalculateLLK0();  // h:334-337
    double Evaluate(const std::vector<double> &v) override;  // h:339-442
  };

  SimplePileupViewer viewer;
  uint32_t NumMarker = 0;
  FullLLKFunc fn;
  std::unordered_map<std::string, std::unordered_map<uint32_t, double>> knownAF;
  double alpha = 0.5;
  std::vector<std::vector<double>> UD;  // [NumMarker][numPC]
  std::vector<std::vector<double>> PC;  // [2][numPC]: PC[0] contaminating, PC[1] intended
  std::vector<double> means;
  BED ChooseBed;
  std::vector<std::pair<std::string, int>> PosVec;

  struct ResolvedMarker {  // h:470-475
    int baseInfoIndex;
    char altBase;
    double knownAFValue;
  };
  std::vector<ResolvedMarker> resolvedMarkers;

  ContaminationEstimator(int nPC, const char *bedFile, int nThread, double ep);  // cpp:38-51
  // cohort mode: share an already parsed panel (.UD/.mu/.bed) instead of reading the files again
  ContaminationEstimator(int nPC, const ContaminationEstimator &panel, int nThread, double ep);
  ContaminationEstimator(const ContaminationEstimator &) = delete;
  ContaminationEstimator &operator=(const ContaminationEstimator &) = delete;
  ~ContaminationEstimator();

  int ReadSVDMatrix(const std::string &UDpath, const std::string &PCpath, const std::string &Mean);  // cpp:334-340
  int ReadMatrixUD(const std::string &path);   // cpp:342-373
  int ReadChooseBed(const std::string &path);  // cpp:413-438
  int ReadMean(const std::string &path);       // cpp:440-459
  int ReadAF(const std::string &path);         // cpp:461-487
  int ReadPileup(const std::string &pileupFile);  // cpp:495-499
  // The same stage on the device (include/vb2_llk.h, vb2_ingest_*): the text is parsed, joined with the panel and --
  // in CreateEngines -- flattened by the GPU; the host only sees the per-row depths its sanity check needs.  Returns
  // false when the text needs the host reader (its parsing quirks, --KnownAF, marker shards): call ReadPileup then.
  bool ReadPileupOnDevice(const std::string &pileupFile, int device);
  vb2_panel *DevicePanel(int device);   // the panel on the device (created on first use, shared by a cohort)
  bool IsSanityCheckOK();                      // cpp:543-587
  void BuildResolvedMarkers();                 // cpp:67-86
  int OptimizeLLK(const std::string &OutputPrefix);  // cpp:88-190
  double RunMinimizer(AmoebaMinimizer &m, const std::vector<double> &startingPoint);  // Reset + Minimize(epsilon)
  bool OptimizeHomoFixedPC(AmoebaMinimizer &m);      // cpp:315-332
  bool OptimizeHomoFixedAlpha(AmoebaMinimizer &m);   // cpp:291-313
  bool OptimizeHomo(AmoebaMinimizer &m);             // cpp:265-289
  bool OptimizeHeterFixedPC(AmoebaMinimizer &m);     // cpp:261-263
  bool OptimizeHeterFixedAlpha(AmoebaMinimizer &m);  // cpp:228-259
  bool OptimizeHeter(AmoebaMinimizer &m);            // cpp:192-226

  // Flatten resolvedMarkers + viewer into the C ABI descriptor and create one engine context per GPU.
  void CreateEngines();
  void DestroyEngines();
  std::vector<vb2_llk_ctx *> engines;
  vb2_panel *devPanel = nullptr;        // owned unless borrowed from the cohort's panel estimator
  bool devPanelOwned = false;
  int devPanelDevice = -1;
  const ContaminationEstimator *panelOwner = nullptr;  // cohort: the estimator whose panel this one shares
  vb2_ingest *devIngest = nullptr;      // a pileup parsed on the device, waiting to be flattened
  const int32_t *devRowDepth = nullptr; // per panel row: kept bases on its pileup line, -1 = no line
  std::function<void()> onEnginesReady;  // called once the sample is resident (cohort: frees a flatten slot)
  double engineSeconds = 0;  // wall time spent inside ComputeMixLLKs
  long deviceSimplexEvals = 0;  // evaluations made by searches that ran on the device
};

}  // namespace vb2
#endif
