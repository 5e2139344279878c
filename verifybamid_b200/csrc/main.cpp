// main.cpp -- command line of the B200 contamination-likelihood engine, flag-compatible with the
// reference for the likelihood path (main.cpp:56-414 of the reference): same option names, same
// defaults, same <out>.selfSM / <out>.Ancestry / <out>.Pileup files, same stdout lines.
//   --BamFile needs htslib (absent here): give --PileupFile instead.
//   --RefVCF (SVD panel construction): svd_panel.cpp + libvb2svd.so (plain or gzip VCF; the decomposition runs on the device).
// Engine-only options: --NumGPU n (marker shards over devices 0..n-1), --Device d, --PanelFP64,
// --PileupList file (cohort mode: many samples on one panel, evaluated in lock-step, see cohort.h).
#include <chrono>
#include <condition_variable>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "cohort.h"
#include "estimator.h"
#include "svd_panel.h"

using namespace vb2;

namespace {

struct Options {
  std::string UDPath = "Empty", MeanPath = "Empty", BedPath = "Empty", BamFile = "Empty", RefPath = "Empty";
  std::string outputPrefix = "result", PileupFile = "Empty", SVDPrefix = "Empty", knownAF = "Empty";
  std::string RefVCF = "Empty", fixPC = "Empty", PileupList = "Empty";
  double fixAlpha = -1., epsilon = 1e-8;  // main.cpp:76
  bool withinAncestry = false, outputPileup = false, verbose = false, disableSanityCheck = false;
  int seed = 12345, nPC = 2, nthread = 4;  // main.cpp:79
  int numGPU = 1, device = 0;
  bool panelFp64 = false;
  // --RefVCF (main.cpp:63-73)
  int numSVDPCs = 10;
  bool skipMinSampleCountCheck = false, gramSVD = false;
  std::string includeChrStr = "1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,"
                              "chr1,chr2,chr3,chr4,chr5,chr6,chr7,chr8,chr9,chr10,"
                              "chr11,chr12,chr13,chr14,chr15,chr16,chr17,chr18,chr19,"
                              "chr20,chr21,chr22";
};

bool ieq(const std::string &a, const char *b) {
  size_t n = strlen(b);
  if (a.size() != n) return false;
  for (size_t i = 0; i < n; ++i)
    if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
  return true;
}

void usage() {
  fprintf(stderr,
          "Options (likelihood path of VerifyBamID2):\n"
          "  --SVDPrefix [String]   SVD files prefix (.UD, .mu, .bed)            | --UDPath --MeanPath --BedPath\n"
          "  --PileupFile [String]  pileup of the sample (samtools mpileup format)\n"
          "  --Reference [String]   reference FASTA (required by the reference CLI; unused with --PileupFile)\n"
          "  --Output [String]      prefix of output files [result]\n"
          "  --NumPC [Int] (2)  --NumThread [Int] (4, host side only)  --Seed [Int] (ignored)  --Epsilon [Double] (1e-8)\n"
          "  --WithinAncestry  --FixPC a:b:...  --FixAlpha x  --KnownAF file  --DisableSanityCheck\n"
          "  --OutputPileup  --Verbose\n"
          "  --NumGPU [Int] (1)  --Device [Int] (0)  --PanelFP64     (engine options)\n"
          "  --PileupList [String]  cohort mode: one '<pileup> [<output prefix>]' per line, all samples on the same\n"
          "                         panel, evaluated in lock-step (one launch per simplex step for the whole cohort;\n"
          "                         with --NumGPU n the samples are spread over n devices)\n"
          "  --RefVCF [String]      build the panel instead: reference VCF (plain or .gz) -> [RefVCF].UD/.mu/.bed/.V,\n"
          "                         SVD on the GPU;  --NumSVDPCs [Int] (10)  --SkipMinSampleCountCheck  --GramSVD\n"
          "                         --IncludeChr a,b,... (default: autosomes 1..22 / chr1..chr22)\n");
}

// libStatGen-style long options: "--Name value", booleans are presence flags (params.cpp:114-185)
bool parse(int argc, char **argv, Options &o) {
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a.size() < 3 || a[0] != '-' || a[1] != '-') {
      fprintf(stderr, "Command line parameter %s (#%d) ignored\n", a.c_str(), i);
      continue;
    }
    std::string name = a.substr(2);
    auto value = [&](std::string &dst) {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for --%s\n", name.c_str()); exit(1); }
      dst = argv[++i];
    };
    auto ivalue = [&](int &dst) { std::string s; value(s); dst = atoi(s.c_str()); };
    auto dvalue = [&](double &dst) { std::string s; value(s); dst = atof(s.c_str()); };
    if (ieq(name, "help")) { usage(); exit(1); }
    else if (ieq(name, "BamFile")) value(o.BamFile);
    else if (ieq(name, "PileupFile")) value(o.PileupFile);
    else if (ieq(name, "Reference")) value(o.RefPath);
    else if (ieq(name, "SVDPrefix")) value(o.SVDPrefix);
    else if (ieq(name, "Output")) value(o.outputPrefix);
    else if (ieq(name, "WithinAncestry")) o.withinAncestry = true;
    else if (ieq(name, "DisableSanityCheck")) o.disableSanityCheck = true;
    else if (ieq(name, "NumPC")) ivalue(o.nPC);
    else if (ieq(name, "FixPC")) value(o.fixPC);
    else if (ieq(name, "FixAlpha")) dvalue(o.fixAlpha);
    else if (ieq(name, "KnownAF")) value(o.knownAF);
    else if (ieq(name, "NumThread")) ivalue(o.nthread);
    else if (ieq(name, "Seed")) ivalue(o.seed);
    else if (ieq(name, "Epsilon")) dvalue(o.epsilon);
    else if (ieq(name, "OutputPileup")) o.outputPileup = true;
    else if (ieq(name, "Verbose")) o.verbose = true;
    else if (ieq(name, "RefVCF")) value(o.RefVCF);
    else if (ieq(name, "NumSVDPCs")) ivalue(o.numSVDPCs);
    else if (ieq(name, "SkipMinSampleCountCheck")) o.skipMinSampleCountCheck = true;
    else if (ieq(name, "GramSVD")) o.gramSVD = true;
    else if (ieq(name, "IncludeChr")) value(o.includeChrStr);
    else if (ieq(name, "UDPath")) value(o.UDPath);
    else if (ieq(name, "MeanPath")) value(o.MeanPath);
    else if (ieq(name, "BedPath")) value(o.BedPath);
    else if (ieq(name, "NumGPU")) ivalue(o.numGPU);
    else if (ieq(name, "Device")) ivalue(o.device);
    else if (ieq(name, "PanelFP64")) o.panelFp64 = true;
    else if (ieq(name, "PileupList")) value(o.PileupList);
    else fprintf(stderr, "Command line parameter %s (#%d) ignored\n", a.c_str(), i);
  }
  return true;
}

struct PhaseTimer {  // main.cpp:40-54
  std::string name;
  std::chrono::steady_clock::time_point start;
  explicit PhaseTimer(const std::string &n) : name(n), start(std::chrono::steady_clock::now()) {
    notice("Starting phase: %s", name.c_str());
  }
  ~PhaseTimer() {
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    notice("Finished phase: %s  [%.3f seconds]", name.c_str(), secs);
  }
};

// main.cpp:285-319: model flags
void configure_model(ContaminationEstimator &Estimator, const Options &o, bool chatty) {
  Estimator.verbose = o.verbose;
  Estimator.seed = o.seed;
  Estimator.isHeter = !o.withinAncestry;
  Estimator.isSanityCheckDisabled = o.disableSanityCheck;
  Estimator.panelFp64 = o.panelFp64;
  if (o.fixPC != "Empty") {
    if (chatty) {
      notice("you specified --fixPC, this will overide dynamic estimation of PCs");
      notice("parsing the PCs");
    }
    std::stringstream ss(o.fixPC);
    std::string token;
    std::vector<double> tmpPC;
    while (std::getline(ss, token, ':')) tmpPC.push_back(atof(token.c_str()));
    if ((int)tmpPC.size() > o.nPC && chatty)
      warning("parameter --fixPC provided larger dimension than parameter --numPC(default value 2) and hence will be truncated");
    if ((int)tmpPC.size() < o.nPC)
      error("parameter --fixPC provided smaller dimension than parameter --numPC(default value 2)");
    for (int i = 0; i < o.nPC; ++i) Estimator.PC[1][i] = tmpPC[i];
    Estimator.isPCFixed = true;
  } else if (fabs(o.fixAlpha + 1.) > std::numeric_limits<double>::epsilon()) {
    if (chatty) notice("you specified --fixAlpha, this will overide dynamic estimation of alpha");
    Estimator.alpha = o.fixAlpha;
    Estimator.isAlphaFixed = true;
  }
  if (o.knownAF != "Empty") {
    Estimator.isAFknown = true;
    Estimator.isPCFixed = true;
    Estimator.isHeter = false;  // under --knownAF we assume the WithinAncestry model
    Estimator.ReadAF(o.knownAF);
  }
}

// vb1-compatible result, main.cpp:386-411
void write_selfsm(ContaminationEstimator &Estimator, const std::string &outputPrefix) {
  const char *headers =
      "#SEQ_ID\tRG\tCHIP_ID\t#SNPS\t#READS\tAVG_DP\tFREEMIX\tFREELK1\tFREELK0\tFREE_RH\tFREE_RA\tCHIPMIX\tCHIPLK1\tCHIPLK0\tCHIP_RH\tCHIP_RA\tDPREF\tRDPHET\tRDPALT";
  std::string fileName(outputPrefix + ".selfSM");
  std::ofstream fout(fileName);
  if (!fout.is_open()) error("Open file %s failed!", fileName.c_str());
  fout << headers << std::endl;
  fout << Estimator.viewer.SEQ_SM << "\tNA\tNA\t" << Estimator.NumMarker << "\t";
  if (Estimator.isPileupInput) fout << "NA";
  else fout << Estimator.viewer.numBases;
  fout << "\t" << Estimator.viewer.avgDepth << "\t"
       << ((Estimator.fn.globalAlpha < 0.5) ? Estimator.fn.globalAlpha : (1.f - Estimator.fn.globalAlpha)) << "\t"
       << -Estimator.fn.llk1 << "\t" << -Estimator.fn.llk0 << "\t"
       << "NA\tNA\t"
       << "NA\tNA\tNA\tNA\tNA\t"
       << "NA\tNA\tNA" << std::endl;
  fout.close();
  if (!fout) error("Errors detected when writing to file %s !", fileName.c_str());
}

// Cohort mode (--PileupList): every sample optimises on its own host thread with the reference's sequential
// Nelder-Mead; one coordinator per GPU turns the samples' concurrent likelihood requests into one launch.
int run_cohort(const Options &o, const std::string &UDPath, const std::string &PCPath, const std::string &MeanPath,
               const std::string &BedPath) {
  struct Sample {
    std::string pileup, prefix, err;
    std::unique_ptr<ContaminationEstimator> E;
    int device = 0, local = 0;
    bool ok = false;
  };
  std::vector<Sample> samples;
  {
    std::ifstream fin(o.PileupList);
    if (!fin.is_open()) error("Open file %s failed!", o.PileupList.c_str());
    std::string line;
    while (std::getline(fin, line)) {
      std::stringstream ss(line);
      Sample smp;
      if (!(ss >> smp.pileup)) continue;
      if (!(ss >> smp.prefix)) smp.prefix = o.outputPrefix + "." + std::to_string(samples.size());
      samples.push_back(std::move(smp));
    }
  }
  if (samples.empty()) error("--PileupList %s names no pileup", o.PileupList.c_str());
  notice("Cohort mode: %d samples on %d GPU(s)", (int)samples.size(), o.numGPU);

  ContaminationEstimator panel(o.nPC, BedPath.c_str(), o.nthread, o.epsilon);
  {
    PhaseTimer t("Load SVD reference data");
    panel.ReadSVDMatrix(UDPath, PCPath, MeanPath);
  }
  if (o.knownAF == "Empty") panel.DevicePanel(o.device);  // the device-side pileup reader of the samples on that GPU
  std::vector<int> per_device(o.numGPU, 0);
  for (size_t i = 0; i < samples.size(); ++i) {
    samples[i].device = o.device + (int)(i % o.numGPU);
    samples[i].local = per_device[i % o.numGPU]++;
  }
  std::vector<std::unique_ptr<CohortCoordinator>> coord;
  for (int d = 0; d < o.numGPU; ++d) coord.emplace_back(new CohortCoordinator(per_device[d], o.nPC));

  auto t0 = std::chrono::steady_clock::now();
  // every sample gets its thread (the lock-step needs all of them alive), but only as many as there are cores read and
  // flatten at the same time (the flatten itself runs up to eight threads)
  struct Gate {
    std::mutex mu;
    std::condition_variable cv;
    int free_slots;
    void enter() { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return free_slots > 0; }); --free_slots; }
    void leave() { { std::lock_guard<std::mutex> l(mu); ++free_slots; } cv.notify_one(); }
  } gate;
  gate.free_slots = std::max(2, (int)std::thread::hardware_concurrency() / 2);
  std::vector<std::thread> workers, coordinators;
  for (int d = 0; d < o.numGPU; ++d)
    if (per_device[d]) coordinators.emplace_back([&, d]() { coord[d]->Run(); });
  for (size_t i = 0; i < samples.size(); ++i)
    workers.emplace_back([&, i]() {
      Sample &smp = samples[i];
      CohortCoordinator &C = *coord[smp.device - o.device];
      try {
        smp.E.reset(new ContaminationEstimator(o.nPC, panel, o.nthread, o.epsilon));
        ContaminationEstimator &E = *smp.E;
        configure_model(E, o, false);
        E.quiet = true;
        E.numGPU = 1;
        E.firstDevice = smp.device;
        E.cohort = &C;
        E.cohortIndex = smp.local;
        gate.enter();
        struct Leave { Gate &g; bool done = false; void now() { if (!done) { done = true; g.leave(); } } ~Leave() { now(); } } leave{gate};
        if (!E.ReadPileupOnDevice(smp.pileup, smp.device)) E.ReadPileup(smp.pileup);
        if (!o.disableSanityCheck && !E.IsSanityCheckOK())
          throw std::runtime_error("Insufficient Available markers (sanity check)");
        E.onEnginesReady = [&leave]() { leave.now(); };  // (the search itself waits on the GPU, not on a core)
        E.OptimizeLLK(smp.prefix);
        write_selfsm(E, smp.prefix);
        smp.ok = true;
      } catch (std::exception &e) {
        smp.err = e.what();
      }
      C.Finish(smp.local);
    });
  for (auto &w : workers) w.join();
  for (auto &c : coordinators) c.join();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  long launches = 0, evals = 0;
  for (auto &c : coord) { launches += c->launches; evals += c->evaluations; }
  std::cout << "#SAMPLE\tOUTPUT\tFREEMIX\tFREELK1\tFREELK0\tSTATUS" << std::endl;
  int failed = 0;
  for (auto &smp : samples) {
    if (smp.ok) {
      auto &fn = smp.E->fn;
      std::cout << smp.pileup << "\t" << smp.prefix << "\t" << (fn.globalAlpha < 0.5 ? fn.globalAlpha : (1 - fn.globalAlpha))
                << "\t" << -fn.llk1 << "\t" << -fn.llk0 << "\tOK" << std::endl;
    } else {
      ++failed;
      std::cout << smp.pileup << "\t" << smp.prefix << "\tNA\tNA\tNA\tFAILED: " << smp.err << std::endl;
    }
  }
  notice("Cohort: %d samples, %ld likelihood evaluations in %ld launches (%.1f per launch), %.3f s", (int)samples.size(),
         evals, launches, launches ? (double)evals / launches : 0.0, secs);
  return failed ? 1 : 0;
}

int execute(int argc, char **argv) {
  Options o;
  parse(argc, argv, o);

  if (o.RefVCF != "Empty") {  // main.cpp:232-257: SVD on the fly, then done
    notice("Specified --RefVCF reference panel VCF file, doing SVD on the fly...");
    notice("This procedure will generate SVD matrices as [RefVCF path].UD and [RefVCF path].mu");
    notice("You may specify --SVDPrefix [RefVCF path](or --UDPath [RefVCF path].UD and --MeanPath [RefVCF path].mu) in future use");
    std::unordered_set<std::string> includeChrSet;
    {
      std::stringstream ss(o.includeChrStr);
      std::string token;
      while (std::getline(ss, token, ','))
        if (!token.empty()) includeChrSet.insert(token);
    }
    notice("--IncludeChr: filtering to %d chromosome name(s)", (int)includeChrSet.size());
    SVDcalculator calculator;
    calculator.ProcessRefVCF(o.RefVCF, includeChrSet, o.skipMinSampleCountCheck, o.numSVDPCs, o.gramSVD, o.device);
    notice("Success!");
    return 0;
  }
  if (o.SVDPrefix == "Empty") {  // main.cpp:214-231
    if (o.UDPath == "Empty") error("--UDPath is required when --RefVCF is absent");
    if (o.MeanPath == "Empty") error("--MeanPath is required when --RefVCF is absent");
    if (o.BedPath == "Empty") error("--BedPath is required when --RefVCF is absent");
  } else {
    o.UDPath = o.SVDPrefix + ".UD";
    o.MeanPath = o.SVDPrefix + ".mu";
    o.BedPath = o.SVDPrefix + ".bed";
  }
  std::string PCPath = o.UDPath.substr(0, o.UDPath.size() - 3) + ".V";
  if (o.RefPath == "Empty") error("--Reference is required");  // main.cpp:263-266
  if (o.BamFile != "Empty") {
    error("--BamFile needs htslib, which this build does not link. Produce a pileup "
          "(samtools mpileup, or --OutputPileup of the reference) and pass it with --PileupFile");
  } else if (o.PileupFile == "Empty" && o.PileupList == "Empty") {
    error("--BamFile or --PileupFile is required");
  }
  if (o.numGPU < 1) error("--NumGPU must be at least 1");

  // CUDA context creation and kernel loading take 0.3-0.5 s in a fresh process: do them on helper threads
  // while this thread reads the panel and the pileup.  Failures are reported when the engine is created.
  std::vector<std::thread> warm;
  for (int g = 0; g < o.numGPU; ++g) warm.emplace_back([=]() { vb2_llk_warmup(o.device + g); });
  struct Joiner {
    std::vector<std::thread> &t;
    ~Joiner() { for (auto &x : t) if (x.joinable()) x.join(); }
  } joiner{warm};

  if (o.PileupList != "Empty") {
    for (auto &x : warm) x.join();
    return run_cohort(o, o.UDPath, PCPath, o.MeanPath, o.BedPath);
  }

  // main.cpp:283-319
  ContaminationEstimator Estimator(o.nPC, o.BedPath.c_str(), o.nthread, o.epsilon);
  Estimator.numGPU = o.numGPU;
  Estimator.firstDevice = o.device;
  configure_model(Estimator, o, true);
  {
    PhaseTimer t("Load SVD reference data");
    Estimator.ReadSVDMatrix(o.UDPath, PCPath, o.MeanPath);
  }
  {
    PhaseTimer t("Read pileup");
    // The device reader needs the CUDA context; --OutputPileup needs the reads on the host.
    bool on_device = false;
    if (!o.outputPileup && o.numGPU == 1) {
      for (auto &x : warm) if (x.joinable()) x.join();
      on_device = Estimator.ReadPileupOnDevice(o.PileupFile, o.device);
    }
    if (!on_device) Estimator.ReadPileup(o.PileupFile);
  }

  if (o.outputPileup) {  // main.cpp:336-369
    std::string fileName(o.outputPrefix + ".Pileup");
    std::ofstream fout(fileName);
    if (!fout.is_open()) error("Open file %s failed!", fileName.c_str());
    for (auto &item : Estimator.PosVec) {
      auto chrIt = Estimator.viewer.posIndex.find(item.first);
      if (chrIt == Estimator.viewer.posIndex.end()) continue;
      auto posIt = chrIt->second.find(item.second);
      if (posIt == chrIt->second.end()) continue;
      const int32_t b = posIt->second;
      const size_t depth = Estimator.viewer.DepthOf(b);
      if (depth > 0) {
        fout << item.first << "\t" << item.second << "\t" << Estimator.ChooseBed[item.first][item.second].first << "\t"
             << depth << "\t";
        fout.write(Estimator.viewer.bases.data() + Estimator.viewer.infoOffset[b], (std::streamsize)depth);
        fout << "\t";
        fout.write(Estimator.viewer.quals.data() + Estimator.viewer.infoOffset[b], (std::streamsize)depth);
        fout << std::endl;
      }
    }
    fout.close();
    if (!fout) error("Errors detected when writing to file %s !", fileName.c_str());
  }

  if (!o.disableSanityCheck) {  // main.cpp:371-379
    PhaseTimer t("Marker sanity check");
    if (Estimator.IsSanityCheckOK()) notice("Passing Marker Sanity Check...");
    else {
      warning("Insufficient Available markers, check input bam depth distribution in output pileup file after specifying --OutputPileup");
      exit(EXIT_FAILURE);
    }
  }
  for (auto &x : warm) if (x.joinable()) x.join();
  {
    PhaseTimer t("Optimize likelihood");
    Estimator.OptimizeLLK(o.outputPrefix);
  }
  write_selfsm(Estimator, o.outputPrefix);
  notice("Success!");
  if (getenv("VB2_CLI_TIMING")) {
    auto t0 = std::chrono::steady_clock::now();
    Estimator.DestroyEngines();
    notice("engine teardown %.3f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  fprintf(stderr, "VerifyBamID2 contamination-likelihood engine for NVIDIA B200 (sm_100a).\n");
  fprintf(stderr, " Same model, options and outputs as VerifyBamID2 (Zhang & Kang) on the --PileupFile path.\n\n");
  const auto t_main = std::chrono::steady_clock::now();
  struct AtExit {
    std::chrono::steady_clock::time_point t0;
    ~AtExit() {
      if (getenv("VB2_CLI_TIMING"))
        fprintf(stderr, "NOTICE - main() returns %.3f s after it was entered\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  } at_exit{t_main};
  try {
    const int rc = execute(argc, argv);
    // Everything is written and closed.  Leave without unwinding the CUDA context: its teardown at exit costs
    // 0.3-0.8 s on this part (more than the whole estimate) and the driver reclaims everything with the process.
    if (rc == 0 && !getenv("VB2_CLI_SLOW_EXIT")) {
      std::cout.flush();
      std::cerr.flush();
      fflush(stdout);
      fflush(stderr);
      at_exit.~AtExit();
      _exit(0);
    }
    return rc;
  } catch (std::exception &e) {  // main.cpp:438-455
    std::string errorMsg = "Exiting due to ERROR:\n\t";
    errorMsg += e.what();
    std::cerr << errorMsg << std::endl;
    return -1;
  }
}
