// estimator.cpp -- see estimator.h.  Host logic of the likelihood path with the reference's control
// flow (ContaminationEstimator.cpp / ContaminationEstimator.h); the per-marker, per-read arithmetic
// lives in llk_engine.cu and is reached only through the C ABI.
#include "estimator.h"

#include "cohort.h"

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace vb2 {

// ---- statgen/Error.cpp:26-79 ---------------------------------------------------------------------
void notice(const char *msg, ...) {
  va_list ap;
  va_start(ap, msg);
  fprintf(stderr, "NOTICE - ");
  vfprintf(stderr, msg, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
void warning(const char *msg, ...) {
  va_list ap;
  va_start(ap, msg);
  fprintf(stderr, "\n\aWARNING - \n");
  vfprintf(stderr, msg, ap);
  fprintf(stderr, "\n");
  va_end(ap);
}
void error(const char *msg, ...) {
  va_list ap;
  va_start(ap, msg);
  fprintf(stderr, "\nFATAL ERROR - \n");
  vfprintf(stderr, msg, ap);
  fprintf(stderr, "\n\n");
  va_end(ap);
  throw std::runtime_error("FATAL ERROR");  // the reference throws pexception here
}

namespace {

// ContaminationEstimator.cpp:10-24
struct PhaseTimer {
  std::string name;
  std::chrono::steady_clock::time_point start;
  explicit PhaseTimer(const std::string &phaseName) : name(phaseName), start(std::chrono::steady_clock::now()) {
    notice("  Starting phase: %s", name.c_str());
  }
  ~PhaseTimer() {
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    notice("  Finished phase: %s  [%.3f seconds]", name.c_str(), secs);
  }
};

// InputFile::readLine returns -1 when EOF arrives before '\n' (statgen/InputFile.h:302-309), so the
// reference's `while (fin.readLine(line)==0)` loops drop an unterminated last line.  Same here.
bool read_terminated_lines(const std::string &path, std::vector<std::string> &lines) {
  std::ifstream fin(path, std::ios::binary);
  if (!fin.is_open()) return false;
  std::string all((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
  size_t beg = 0;
  for (;;) {
    size_t nl = all.find('\n', beg);
    if (nl == std::string::npos) break;
    lines.emplace_back(all, beg, nl - beg);
    beg = nl + 1;
  }
  return true;
}

[[noreturn]] void open_failed(const std::string &path) {
  std::cerr << "Open file:" << path << "\t failed, exit!";
  exit(EXIT_FAILURE);
}

}  // namespace

// ---- FullLLKFunc ---------------------------------------------------------------------------------
double ContaminationEstimator::FullLLKFunc::InvLogit(double x) {
  double e = exp(x);
  return e / (1. + e);
}
double ContaminationEstimator::FullLLKFunc::Logit(double x) { return log(x / (1. - x)); }

double ContaminationEstimator::FullLLKFunc::ComputeMixLLKs(const std::vector<double> &tPC1,
                                                           const std::vector<double> &tPC2, double alpha) {
  ContaminationEstimator &E = *ptr;
  if (E.engines.empty()) error("ComputeMixLLKs called before the engine was created");
  auto t0 = std::chrono::steady_clock::now();
  ++evalCount;
  if (E.cohort) {  // lock-step with the other samples of the cohort: one launch evaluates all of them
    const double llk = E.cohort->Evaluate(E.cohortIndex, E.engines[0], tPC1.data(), tPC2.data(), alpha);
    E.engineSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return llk;
  }
  // every marker shard (one per GPU) evaluates concurrently; partial sums are added in shard order
  for (vb2_llk_ctx *c : E.engines)
    if (vb2_llk_eval_begin(c, tPC1.data(), tPC2.data(), alpha) != VB2_OK) error("GPU engine: %s", vb2_last_error(c));
  double sumLLK = 0;
  for (vb2_llk_ctx *c : E.engines) {
    double part = 0;
    if (vb2_llk_eval_end(c, &part) != VB2_OK) error("GPU engine: %s", vb2_last_error(c));
    sumLLK += part;
  }
  E.engineSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return sumLLK;
}

int ContaminationEstimator::FullLLKFunc::Initialize() {
  globalPC = fixPC = globalPC2 = fixPC2 = ptr->PC[1];  // only the intended sample has pre-defined PCs
  globalAlpha = fixAlpha = ptr->alpha;
  llk1 = (0 - ComputeMixLLKs(fixPC, fixPC2, fixAlpha));
  for (int k = 0; k < ptr->numPC; ++k) ptr->PC[0][k] = 0.01;
  for (int k = 0; k < ptr->numPC; ++k) ptr->PC[1][k] = 0.01;
  ptr->alpha = 0.03;
  return 0;
}

int ContaminationEstimator::FullLLKFunc::CalculateLLK0() {
  llk0 = (0 - ComputeMixLLKs(globalPC, globalPC, 0));
  return 0;
}

double ContaminationEstimator::FullLLKFunc::Evaluate(const std::vector<double> &v) {
  double smLLK = 0;
  const int numPC = ptr->numPC;
  if (!ptr->isHeter) {
    if (ptr->isPCFixed) {
      double tmpAlpha = InvLogit(v[0]);
      smLLK = 0 - ComputeMixLLKs(fixPC, fixPC2, tmpAlpha);
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalAlpha = tmpAlpha;
      }
    } else if (ptr->isAlphaFixed) {
      std::vector<double> tmpPC(v.begin(), v.begin() + numPC);
      smLLK = 0 - ComputeMixLLKs(tmpPC, tmpPC, fixAlpha);
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalPC = tmpPC;
        globalPC2 = tmpPC;
      }
    } else {
      std::vector<double> tmpPC(v.begin(), v.begin() + numPC);
      double tmpAlpha = InvLogit(v[numPC]);
      smLLK = 0 - ComputeMixLLKs(tmpPC, tmpPC, tmpAlpha);
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalPC = tmpPC;
        globalPC2 = tmpPC;
        globalAlpha = tmpAlpha;
      }
    }
  } else {  // contamination source from a different population
    if (ptr->isPCFixed) {  // only fixed for the intended sample
      std::vector<double> tmpPC(v.begin(), v.begin() + numPC);
      double tmpAlpha = InvLogit(v[numPC]);
      smLLK = 0 - ComputeMixLLKs(tmpPC, fixPC2, tmpAlpha);
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalPC = tmpPC;
        globalAlpha = tmpAlpha;
      }
    } else if (ptr->isAlphaFixed) {
      std::vector<double> tmpPC(numPC, 0.), tmpPC2(numPC, 0.);
      for (int k = 0; k < (int)v.size(); ++k) {
        if (k < numPC) tmpPC[k] = v[k];
        else if (k < numPC * 2) tmpPC2[k - numPC] = v[k];
        else error("Simplex Vector dimension error!");
      }
      smLLK = 0 - ComputeMixLLKs(tmpPC, tmpPC2, fixAlpha);
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalPC = tmpPC;
        globalPC2 = tmpPC2;
      }
    } else {
      std::vector<double> tmpPC(numPC, 0.), tmpPC2(numPC, 0.);
      double tmpAlpha = 0.;
      for (int k = 0; k < (int)v.size(); ++k) {
        if (k < numPC) tmpPC[k] = v[k];
        else if (k < numPC * 2) tmpPC2[k - numPC] = v[k];
        else if (k == numPC * 2) tmpAlpha = InvLogit(v[k]);
        else error("Simplex Vector dimension error!");
      }
      smLLK = (0 - ComputeMixLLKs(tmpPC, tmpPC2, tmpAlpha));
      if (smLLK < llk1) {
        llk1 = smLLK;
        globalPC = tmpPC;
        globalPC2 = tmpPC2;
        globalAlpha = tmpAlpha;
      }
    }
  }
  if (ptr->verbose)
    notice("ContaminatingSamplePC1:%f\tContaminatingSamplePC2:%f\tIntendedSamplePC1:%f\tIntendedSamplePC2:%f\tFREEMIX(Alpha):%f\tllk:%f",
           globalPC[0], numPC > 1 ? globalPC[1] : 0., globalPC2[0], numPC > 1 ? globalPC2[1] : 0., globalAlpha, llk1);
  return smLLK;
}

// ---- ContaminationEstimator -------------------------------------------------------------------------
ContaminationEstimator::ContaminationEstimator(int nPC, const char *bedFile, int nThread, double ep)
    : numPC(nPC), numThread(nThread), epsilon(ep), PC(2, std::vector<double>(nPC, 0.)) {
  fn.ptr = this;
  fn.fixPC.assign(nPC, 0.);
  fn.fixPC2.assign(nPC, 0.);
  fn.globalPC = fn.fixPC;
  fn.globalPC2 = fn.fixPC2;
  std::cerr << "Initialize from FullLLKFunc(int dim, ContaminationEstimator* contPtr)" << std::endl;  // h:114
  ReadChooseBed(std::string(bedFile));
  alpha = 0.5;
  NumMarker = 0;
}

ContaminationEstimator::ContaminationEstimator(int nPC, const ContaminationEstimator &panel, int nThread, double ep)
    : numPC(nPC), numThread(nThread), epsilon(ep), PC(2, std::vector<double>(nPC, 0.)) {
  fn.ptr = this;
  fn.fixPC.assign(nPC, 0.);
  fn.fixPC2.assign(nPC, 0.);
  fn.globalPC = fn.fixPC;
  fn.globalPC2 = fn.fixPC2;
  alpha = 0.5;
  NumMarker = panel.NumMarker;
  UD = panel.UD;
  means = panel.means;
  ChooseBed = panel.ChooseBed;
  PosVec = panel.PosVec;
  panelOwner = &panel;
}

ContaminationEstimator::~ContaminationEstimator() {
  DestroyEngines();
  if (devIngest) vb2_ingest_destroy(devIngest);
  if (devPanel && devPanelOwned) vb2_panel_destroy(devPanel);
}

vb2_panel *ContaminationEstimator::DevicePanel(int device) {
  if (devPanel) return devPanel;
  if (PosVec.size() < NumMarker || means.size() < NumMarker) return nullptr;
  // chromosome names -> small integers; per row the position and the ALT base as ChooseBed resolves it (cpp:80)
  std::vector<std::string> names;
  std::unordered_map<std::string, uint16_t> id;
  std::vector<uint16_t> chrom(NumMarker);
  std::vector<int32_t> pos(NumMarker);
  std::vector<char> alt(NumMarker);
  std::vector<double> ud((size_t)NumMarker * numPC);
  for (size_t i = 0; i < NumMarker; ++i) {
    const std::string &chr = PosVec[i].first;
    auto it = id.find(chr);
    if (it == id.end()) {
      if (names.size() >= 65535) return nullptr;
      it = id.emplace(chr, (uint16_t)names.size()).first;
      names.push_back(chr);
    }
    chrom[i] = it->second;
    pos[i] = PosVec[i].second;
    alt[i] = ChooseBed[chr][PosVec[i].second].second;
    for (int k = 0; k < numPC; ++k) ud[i * numPC + k] = UD[i][k];
  }
  std::string blob;
  for (const std::string &n : names) { blob += n; blob += '\0'; }
  vb2_panel_desc d;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.n_marker = NumMarker;
  d.n_pc = (uint32_t)numPC;
  d.ud_stride = (uint32_t)numPC;
  d.ud = ud.data();
  d.means = means.data();
  d.chrom_id = chrom.data();
  d.pos = pos.data();
  d.alt_base = alt.data();
  d.chrom_names = blob.data();
  d.n_chrom = (uint32_t)names.size();
  d.device = device;
  if (vb2_panel_create(&d, &devPanel) != VB2_OK) {
    devPanel = nullptr;
    return nullptr;
  }
  devPanelOwned = true;
  devPanelDevice = device;
  return devPanel;
}

bool ContaminationEstimator::ReadPileupOnDevice(const std::string &pileupFile, int device) {
  if (isAFknown || numGPU != 1 || getenv("VB2_HOST_INGEST")) return false;
  // (a cohort shares the panel its panel estimator put on ONE device; samples on other devices use the host reader)
  vb2_panel *panel = panelOwner ? (panelOwner->devPanelDevice == device ? panelOwner->devPanel : nullptr) : DevicePanel(device);
  if (!panel) return false;
  std::ifstream fin(pileupFile, std::ios::binary);
  if (!fin.is_open()) error("open file %s failed!", pileupFile.c_str());
  std::string text((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
  vb2_ingest_info info;
  memset(&info, 0, sizeof(info));
  info.struct_size = sizeof(info);
  if (devIngest) { vb2_ingest_destroy(devIngest); devIngest = nullptr; }
  const int rc = vb2_ingest_parse(panel, text.data(), text.size(), &devIngest, &info);
  if (rc == VB2_ERR_UNSUPPORTED) {
    notice("pileup text is read by the host (%s)", vb2_last_error(nullptr));
    return false;
  }
  if (rc != VB2_OK) error("GPU pileup ingest: %s", vb2_last_error(nullptr));
  viewer = SimplePileupViewer();
  viewer.numBases = (long)info.num_bases;            // cpp:826
  viewer.effectiveNumSite = (int)info.n_matched;     // cpp:829
  viewer.avgDepth = (double)viewer.numBases / viewer.GetNumMarker();  // cpp:831
  devRowDepth = info.row_depth;
  isPileupInput = true;
  return true;
}

void ContaminationEstimator::BuildResolvedMarkers() {
  resolvedMarkers.resize(NumMarker);
  for (size_t i = 0; i < NumMarker; ++i) {
    const std::string &chr = PosVec[i].first;
    int pos = PosVec[i].second;
    ResolvedMarker &rm = resolvedMarkers[i];
    rm.altBase = 0;
    rm.knownAFValue = 0.0;
    auto chrIt = viewer.posIndex.find(chr);
    if (chrIt == viewer.posIndex.end()) { rm.baseInfoIndex = -1; continue; }
    auto posIt = chrIt->second.find(pos);
    if (posIt == chrIt->second.end()) { rm.baseInfoIndex = -1; continue; }
    rm.baseInfoIndex = posIt->second;
    rm.altBase = ChooseBed[chr][pos].second;
    if (isAFknown) rm.knownAFValue = knownAF[chr][pos];
  }
}

void ContaminationEstimator::CreateEngines() {
  DestroyEngines();
  if (devIngest) {  // the pileup sits on the device already: flatten it there
    vb2_flatten_desc f;
    memset(&f, 0, sizeof(f));
    f.struct_size = sizeof(f);
    f.device = firstDevice;
    f.panel_dtype = panelFp64 ? VB2_PANEL_FP64 : VB2_PANEL_FP32;
    if (cohort) f.flags |= VB2_FLAG_BATCHED;
    f.sanity_disabled = isSanityCheckDisabled ? 1 : 0;
    f.avg_depth = viewer.avgDepth;
    f.sd_depth = viewer.sdDepth;
    vb2_llk_ctx *c = nullptr;
    const int rc = vb2_ingest_flatten(devIngest, &f, &c);
    vb2_ingest_destroy(devIngest);
    devIngest = nullptr;
    devRowDepth = nullptr;
    if (rc != VB2_OK) error("cannot flatten the pileup on device %d: %s", firstDevice, vb2_last_error(nullptr));
    engines.push_back(c);
    vb2_llk_info info;
    info.struct_size = sizeof(info);
    if (vb2_llk_get_info(c, &info) == VB2_OK)
      notice("GPU likelihood engine: 1 device(s), %llu markers / %llu reads resident in HBM (parsed and flattened on the device)",
             (unsigned long long)info.markers_used, (unsigned long long)info.reads_used);
    if (!cohort && !getenv("VB2_NO_SESSION") && vb2_llk_session_begin(c) == VB2_OK)
      notice("evaluation session: resident kernel on 1 device(s)");
    return;
  }
  if (PosVec.size() < NumMarker || means.size() < NumMarker)
    error("SVD files disagree: %u rows in .UD, %d in .bed, %d in .mu", NumMarker, (int)PosVec.size(), (int)means.size());
  std::vector<double> ud((size_t)NumMarker * numPC);
  std::vector<int32_t> idx(NumMarker);
  std::vector<char> alt(NumMarker);
  std::vector<double> kaf;
  if (isAFknown) kaf.resize(NumMarker);
  for (size_t i = 0; i < NumMarker; ++i) {
    for (int k = 0; k < numPC; ++k) ud[i * numPC + k] = UD[i][k];
    idx[i] = resolvedMarkers[i].baseInfoIndex;
    alt[i] = resolvedMarkers[i].altBase;
    if (isAFknown) kaf[i] = resolvedMarkers[i].knownAFValue;
  }
  vb2_llk_desc d;
  memset(&d, 0, sizeof(d));
  d.struct_size = sizeof(d);
  d.n_marker = NumMarker;
  d.n_pc = (uint32_t)numPC;
  d.ud_stride = (uint32_t)numPC;
  d.ud = ud.data();
  d.means = means.data();
  d.base_info_index = idx.data();
  d.alt_base = alt.data();
  d.known_af = isAFknown ? kaf.data() : nullptr;
  d.info_offset = viewer.infoOffset.data();
  d.n_info = (int64_t)viewer.NumInfo();
  d.bases = viewer.bases.data();
  d.quals = viewer.quals.data();
  d.sanity_disabled = isSanityCheckDisabled ? 1 : 0;
  d.avg_depth = viewer.avgDepth;
  d.sd_depth = viewer.sdDepth;
  d.panel_dtype = panelFp64 ? VB2_PANEL_FP64 : VB2_PANEL_FP32;
  if (cohort) d.flags |= VB2_FLAG_BATCHED;  // a cohort member is only ever evaluated in the cohort's batched launches
  d.shard_count = (uint32_t)numGPU;
  for (int g = 0; g < numGPU; ++g) {
    d.shard_rank = (uint32_t)g;
    d.device = firstDevice + g;
    vb2_llk_ctx *c = nullptr;
    if (vb2_llk_create(&d, &c) != VB2_OK) {
      std::string why = vb2_last_error(nullptr);
      DestroyEngines();
      error("cannot create the GPU likelihood engine on device %d: %s", d.device, why.c_str());
    }
    engines.push_back(c);
  }
  vb2_llk_info info;
  info.struct_size = sizeof(info);
  uint64_t markers = 0, reads = 0;
  for (vb2_llk_ctx *c : engines)
    if (vb2_llk_get_info(c, &info) == VB2_OK) { markers += info.markers_used; reads += info.reads_used; }
  notice("GPU likelihood engine: %d device(s), %llu markers / %llu reads resident in HBM", numGPU,
         (unsigned long long)markers, (unsigned long long)reads);
  // The simplex search evaluates this one sample several hundred times, each evaluation waiting for the last:
  // keep a resident kernel with the sample in shared memory for the duration (include/vb2_llk.h, evaluation
  // session).  Samples that do not fit on chip simply keep one launch per evaluation; a cohort evaluates all its
  // samples in one launch per step instead.
  if (!cohort && !getenv("VB2_NO_SESSION")) {
    int n_session = 0;
    for (vb2_llk_ctx *c : engines) n_session += vb2_llk_session_begin(c) == VB2_OK;
    if (n_session) notice("evaluation session: resident kernel on %d device(s)", n_session);
  }
}

void ContaminationEstimator::DestroyEngines() {
  for (vb2_llk_ctx *c : engines) vb2_llk_destroy(c);
  engines.clear();
}

int ContaminationEstimator::OptimizeLLK(const std::string &OutputPrefix) {
  AmoebaMinimizer myMinimizer;
  if (!devIngest) BuildResolvedMarkers();
  {
    PhaseTimer t("Flatten pileup into HBM");
    CreateEngines();
  }
  if (onEnginesReady) onEnginesReady();
  {
    PhaseTimer t("Initialize likelihood");
    fn.Initialize();
  }
  if (!isHeter) {
    if (isPCFixed) {
      if (!quiet) std::cout << "Estimation from OptimizeHomoFixedPC:" << std::endl;
      PhaseTimer t("OptimizeHomoFixedPC");
      OptimizeHomoFixedPC(myMinimizer);
    } else if (isAlphaFixed) {
      PhaseTimer t("OptimizeHomoFixedAlpha");
      OptimizeHomoFixedAlpha(myMinimizer);
    } else {
      if (!quiet) std::cout << "Estimation from OptimizeHomo:" << std::endl;
      PhaseTimer t("OptimizeHomo");
      OptimizeHomo(myMinimizer);
    }
  } else {  // contamination source from a different population
    if (isPCFixed) {
      if (!quiet) std::cout << "Estimation from OptimizeHeterFixedPC:" << std::endl;
      PhaseTimer t("OptimizeHeterFixedPC");
      OptimizeHeterFixedPC(myMinimizer);
    } else if (isAlphaFixed) {
      if (!quiet) std::cout << "Estimation from OptimizeHeterFixedAlpha:" << std::endl;
      {
        PhaseTimer t("OptimizeHomoFixedAlpha (initial)");
        isHeter = false;
        OptimizeHomoFixedAlpha(myMinimizer);
        PC[1] = PC[0];
        fn.globalPC2 = fn.globalPC;
        isHeter = true;
      }
      {
        PhaseTimer t("OptimizeHeterFixedAlpha");
        OptimizeHeterFixedAlpha(myMinimizer);
      }
    } else {
      if (!quiet) std::cout << "Estimation from OptimizeHeter:" << std::endl;
      {
        PhaseTimer t("OptimizeHomo (initial)");
        isHeter = false;
        OptimizeHomo(myMinimizer);
        PC[1] = PC[0];
        fn.globalPC2 = fn.globalPC;
        isHeter = true;
      }
      {
        PhaseTimer t("OptimizeHeter");
        OptimizeHeter(myMinimizer);
      }
    }
    if (fn.globalAlpha >= 0.5) {  // cpp:146-149: only PC1 and PC2 are swapped
      std::swap(fn.globalPC[0], fn.globalPC2[0]);
      if (numPC > 1) std::swap(fn.globalPC[1], fn.globalPC2[1]);
    }
  }
  {
    PhaseTimer t("Calculate null-model LLK");
    fn.CalculateLLK0();
  }
  if (!quiet) {
    std::cout << "Contaminating Sample ";
    for (int i = 0; i < numPC; ++i) std::cout << "PC" << i + 1 << ":" << fn.globalPC[i] << "\t";
    std::cout << std::endl;
    std::cout << "Intended Sample ";
    for (int i = 0; i < numPC; ++i) std::cout << "PC" << i + 1 << ":" << fn.globalPC2[i] << "\t";
    std::cout << std::endl;
    std::cout << "FREEMIX(Alpha):" << (fn.globalAlpha < 0.5 ? fn.globalAlpha : (1 - fn.globalAlpha)) << std::endl;
  }

  std::string fileName(OutputPrefix + ".Ancestry");
  std::ofstream fout(fileName);
  if (!fout.is_open()) error("Open file %s failed!", fileName.c_str());
  fout << "PC\tContaminatingSample\tIntendedSample" << std::endl;
  for (int i = 0; i < numPC; ++i) fout << i + 1 << "\t" << fn.globalPC[i] << "\t" << fn.globalPC2[i] << std::endl;
  fout.close();
  if (!fout) error("Errors detected when writing to file %s !", fileName.c_str());
  notice("Likelihood evaluations: %ld, %.3f ms inside the GPU engine (%.1f us per evaluation)", fn.evalCount,
         engineSeconds * 1e3, fn.evalCount ? engineSeconds * 1e6 / fn.evalCount : 0.0);
  if (deviceSimplexEvals) notice("Simplex search on the device: %ld of the evaluations", deviceSimplexEvals);
  return 0;
}

// The simplex search of every Optimize* (ContaminationEstimator.cpp:192-332: Reset, point = start, Minimize(epsilon)).
// With one resident engine it runs next to the kernel (vb2_llk_minimize: AmoebaMinimizer::Minimize on the device,
// Evaluate's unpacking as a table); otherwise -- marker shards on several GPUs, a cohort, --Verbose, a sample that
// does not fit on chip, more than four PCs -- the host drives it evaluation by evaluation, as the reference does.
double ContaminationEstimator::RunMinimizer(AmoebaMinimizer &myMinimizer, const std::vector<double> &startingPoint) {
  const int dim = (int)startingPoint.size();
  myMinimizer.func = &fn;
  myMinimizer.Reset(dim);
  myMinimizer.point = startingPoint;
  if (engines.size() == 1 && !cohort && !verbose && dim <= VB2_MIN_MAX_DIM && numPC <= 4 && !getenv("VB2_HOST_SIMPLEX")) {
    vb2_llk_model M;
    memset(&M, 0, sizeof(M));
    M.struct_size = sizeof(M);
    M.dim = (uint32_t)dim;
    // FullLLKFunc::Evaluate's six branches (h:339-442): which entries of v feed which argument
    const bool pcFree = !isPCFixed, alphaFree = !isAlphaFixed || isPCFixed;  // (isPCFixed is tested first)
    for (int k = 0; k < VB2_MAX_PC; ++k) M.pc1_from[k] = M.pc2_from[k] = -1;
    M.alpha_from = -1;
    M.alpha_fixed = fn.fixAlpha;
    for (int k = 0; k < numPC; ++k) {
      M.pc1_fixed[k] = fn.fixPC[k];
      M.pc2_fixed[k] = fn.fixPC2[k];
    }
    if (!isHeter) {
      if (isPCFixed) M.alpha_from = 0;                                    // v = (logit alpha); PCs fixed
      else {
        for (int k = 0; k < numPC; ++k) M.pc1_from[k] = M.pc2_from[k] = k;  // one set of PCs for both samples
        if (alphaFree) M.alpha_from = numPC;
      }
    } else {
      if (isPCFixed) {                                                    // intended PCs fixed, contaminant's free
        for (int k = 0; k < numPC; ++k) M.pc1_from[k] = k;
        M.alpha_from = numPC;
      } else {
        for (int k = 0; k < numPC; ++k) { M.pc1_from[k] = k; M.pc2_from[k] = numPC + k; }
        if (alphaFree) M.alpha_from = 2 * numPC;
      }
    }
    (void)pcFree;
    vb2_llk_min_result R;
    memset(&R, 0, sizeof(R));
    R.struct_size = sizeof(R);
    auto t0 = std::chrono::steady_clock::now();
    const int rc = vb2_llk_minimize(engines[0], &M, startingPoint.data(), 1.0, epsilon, myMinimizer.cycleMax, fn.llk1, &R);
    if (rc == VB2_OK) {
      engineSeconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      fn.evalCount += (long)R.evals;
      deviceSimplexEvals += (long)R.evals;
      if (R.improved) {  // Evaluate's best-so-far bookkeeping, for the components this model leaves free
        fn.llk1 = R.llk1;
        for (int k = 0; k < numPC; ++k) {
          if (M.pc1_from[k] >= 0) fn.globalPC[k] = R.best_pc_contam[k];
          if (M.pc2_from[k] >= 0) fn.globalPC2[k] = R.best_pc_intended[k];
        }
        if (M.alpha_from >= 0) fn.globalAlpha = R.best_alpha;
      }
      myMinimizer.point.assign(R.point, R.point + dim);
      myMinimizer.cycleCount = (long)R.cycle_count;
      if (!R.converged) {
        fprintf(stderr, "WARNING - Amoeba.Minimize - Couldn't converge in %ld cycles\n", myMinimizer.cycleMax);
        return std::numeric_limits<double>::max();
      }
      return myMinimizer.fmin = R.fmin;
    }
    if (rc != VB2_ERR_INVALID) error("GPU engine: %s", vb2_last_error(engines[0]));
    // (no session on this sample, or a shape the device search does not take: the host drives the simplex)
  }
  return myMinimizer.Minimize(epsilon);
}

bool ContaminationEstimator::OptimizeHeter(AmoebaMinimizer &myMinimizer) {
  std::vector<double> startingPoint(numPC * 2 + 1);
  for (int i = 0; i < numPC * 2; ++i) startingPoint[i] = i < numPC ? PC[0][i] : PC[1][i - numPC];
  startingPoint[numPC * 2] = FullLLKFunc::Logit(alpha);
  if (verbose) {
    std::cerr << "Start point:";
    for (int i = 0; i < numPC * 2; ++i) std::cerr << startingPoint[i] << "\t";
    std::cerr << "and alpha:\t" << alpha << std::endl;
  }
  double ret = RunMinimizer(myMinimizer, startingPoint);
  alpha = FullLLKFunc::InvLogit(myMinimizer.point[numPC * 2]);
  for (int i = 0; i < numPC; ++i) PC[0][i] = myMinimizer.point[i];
  for (int i = numPC; i < numPC * 2; ++i) PC[1][i - numPC] = myMinimizer.point[i];
  return ret != std::numeric_limits<double>::max();
}

bool ContaminationEstimator::OptimizeHeterFixedAlpha(AmoebaMinimizer &myMinimizer) {
  std::vector<double> startingPoint(numPC * 2);
  for (int i = 0; i < numPC * 2; ++i) startingPoint[i] = i < numPC ? PC[0][i] : PC[1][i - numPC];
  if (verbose) {
    std::cerr << "Start point:";
    for (int i = 0; i < numPC * 2; ++i) std::cerr << startingPoint[i] << "\t";
  }
  RunMinimizer(myMinimizer, startingPoint);
  for (int i = 0; i < numPC; ++i) PC[0][i] = myMinimizer.point[i];
  for (int i = numPC; i < numPC * 2; ++i) PC[1][i - numPC] = myMinimizer.point[i];
  return true;  // fixAlpha usually converges well
}

bool ContaminationEstimator::OptimizeHeterFixedPC(AmoebaMinimizer &myMinimizer) { return OptimizeHomo(myMinimizer); }

bool ContaminationEstimator::OptimizeHomo(AmoebaMinimizer &myMinimizer) {
  std::vector<double> startingPoint(numPC + 1);
  for (int i = 0; i < numPC; ++i) startingPoint[i] = PC[0][i];
  startingPoint[numPC] = FullLLKFunc::Logit(alpha);
  if (verbose) {
    std::cerr << "Start point:";
    for (int i = 0; i < numPC; ++i) std::cerr << startingPoint[i] << "\t";
    std::cerr << "and alpha:\t" << alpha << std::endl;
  }
  double ret = RunMinimizer(myMinimizer, startingPoint);
  alpha = FullLLKFunc::InvLogit(myMinimizer.point[numPC]);
  for (int i = 0; i < numPC; ++i) PC[0][i] = myMinimizer.point[i];
  return ret != std::numeric_limits<double>::max();
}

bool ContaminationEstimator::OptimizeHomoFixedAlpha(AmoebaMinimizer &myMinimizer) {
  std::vector<double> startingPoint(numPC);
  for (int i = 0; i < numPC; ++i) startingPoint[i] = PC[0][i];
  if (verbose) {
    std::cerr << "Start point:";
    for (int i = 0; i < numPC; ++i) std::cerr << startingPoint[i] << "\t";
  }
  RunMinimizer(myMinimizer, startingPoint);
  for (int i = 0; i < numPC; ++i) PC[0][i] = myMinimizer.point[i];
  return true;  // fixAlpha usually converges well
}

bool ContaminationEstimator::OptimizeHomoFixedPC(AmoebaMinimizer &myMinimizer) {
  std::vector<double> startingPoint(1);
  startingPoint[0] = FullLLKFunc::Logit(alpha);
  if (verbose) {
    std::cerr << "Start point";
    std::cerr << "alpha:\t" << alpha << std::endl;
  }
  double ret = RunMinimizer(myMinimizer, startingPoint);
  alpha = FullLLKFunc::InvLogit(myMinimizer.point[0]);
  return ret != std::numeric_limits<double>::max();
}

int ContaminationEstimator::ReadSVDMatrix(const std::string &UDpath, const std::string &, const std::string &Mean) {
  ReadMatrixUD(UDpath);
  ReadMean(Mean);
  return 0;
}

int ContaminationEstimator::ReadMatrixUD(const std::string &path) {
  std::vector<std::string> lines;
  if (!read_terminated_lines(path, lines)) open_failed(path);
  std::vector<double> tmpUD(numPC, 0);
  UD.reserve(lines.size());
  for (const std::string &line : lines) {
    std::stringstream ss(line);
    int index = 0;
    while (index < numPC && ss >> tmpUD[index]) index++;
    if (index < numPC) {  // cpp:358-363
      warning("--NumPC should be less than or equal to the number of PCs in SVD files provided by --SVDPrefix! (Expected:%d vs Observed:%d)", numPC, index);
      warning("--NumPC only permits as large as 4 PCs when using SVD files in ${verifybamID}/resource/ directory!");
      warning("You can always prepare you own SVD files with arbitrary number of PCs with --RefVCF enabled.");
      exit(EXIT_FAILURE);
    }
    UD.push_back(tmpUD);
    NumMarker++;
  }
  return 0;
}

int ContaminationEstimator::ReadChooseBed(const std::string &path) {
  std::vector<std::string> lines;
  if (!read_terminated_lines(path, lines)) open_failed(path);
  std::string chr;
  int pos(0);
  char ref(0), alt(0);
  for (const std::string &line : lines) {
    std::stringstream ss(line);
    ss >> chr >> pos >> pos;
    ss >> ref >> alt;  // single chars: ALT "G,T" is read as 'G' (cpp:417,429)
    PosVec.push_back(std::make_pair(chr, pos));
    ChooseBed[chr][pos] = std::make_pair(ref, alt);
  }
  return 0;
}

int ContaminationEstimator::ReadMean(const std::string &path) {
  std::vector<std::string> lines;
  if (!read_terminated_lines(path, lines)) open_failed(path);
  double mu(0);
  std::string snpName;
  for (const std::string &line : lines) {
    std::stringstream ss(line);
    ss >> snpName;
    ss >> mu;
    means.push_back(mu);
  }
  return 0;
}

int ContaminationEstimator::ReadAF(const std::string &path) {
  std::ifstream fin(path);
  std::string line, chr;
  uint32_t pos(0);
  double AF(0);
  char ref(0), alt(0);
  if (!fin.is_open()) open_failed(path);
  while (std::getline(fin, line)) {
    std::stringstream ss(line);
    ss >> chr;
    ss >> pos >> pos;
    ss >> ref >> alt;
    ss >> AF;
    knownAF[chr][pos] = AF;
  }
  return 0;
}

int ContaminationEstimator::ReadPileup(const std::string &pileupFile) {
  viewer = SimplePileupViewer();
  viewer.ReadPileup(pileupFile, ChooseBed);
  isPileupInput = true;
  return 0;
}

bool ContaminationEstimator::IsSanityCheckOK() {
  notice("Number of marker in Reference Matrix:%d", NumMarker);
  notice("Number of marker shared with input file:%d", viewer.GetNumMarker());
  auto depth_at = [&](size_t i, int &depth) -> bool {
    if (devRowDepth) {  // the pileup was parsed on the device: the join is done, the depths are here
      if (devRowDepth[i] < 0) return false;
      depth = devRowDepth[i];
      return true;
    }
    auto chrIt = viewer.posIndex.find(PosVec[i].first);
    if (chrIt == viewer.posIndex.end()) return false;
    auto posIt = chrIt->second.find(PosVec[i].second);
    if (posIt == chrIt->second.end()) return false;
    depth = (int)viewer.DepthOf(posIt->second);
    return true;
  };
  int tmpDepth = 0;
  for (size_t i = 0; i < NumMarker; ++i)
    if (depth_at(i, tmpDepth)) viewer.sdDepth += tmpDepth * tmpDepth;  // (int multiply, as in the reference)
  viewer.sdDepth = sqrt(viewer.sdDepth / viewer.effectiveNumSite - viewer.avgDepth * viewer.avgDepth);
  viewer.effectiveNumSite = 0;
  for (size_t i = 0; i < NumMarker; ++i) {
    if (!depth_at(i, tmpDepth)) continue;
    if (tmpDepth == 0 || tmpDepth < (viewer.avgDepth - 3 * viewer.sdDepth) ||
        tmpDepth > (viewer.avgDepth + 3 * viewer.sdDepth))
      continue;
    viewer.effectiveNumSite++;
  }
  notice("Mean Depth:%f", viewer.avgDepth);
  notice("SD Depth:%f", viewer.sdDepth);
  notice("%d SNP markers remained after sanity check.", viewer.GetNumMarker());
  return viewer.GetNumMarker() > 1000 && viewer.GetNumMarker() > (NumMarker * 0.1);
}

}  // namespace vb2
