// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin command-line driver around the UNMODIFIED reference classes
// (ContaminationEstimator, AmoebaMinimizer, SimplePileupViewer text half), compiled by
// oracle/Makefile straight from /root/reference into oracle/_ref/vb2_ref.  It exists
// only so that tests and bench.py's cpu_baseline / `--impl reference` leg can run the
// reference's own CPU implementation of the hot path.
//
// It restates the pileup-input slice of the reference CLI flow
// (main.cpp:283-333 estimator set-up + load, :371-379 sanity check, :381-384 optimise,
// :386-411 .selfSM) without the BAM branch (needs htslib) and without PhoneHome.
// Extra, driver-only switches:
//   --EvalPoints <file>   one "pc_contam[k] pc_intended[k] alpha" per line; prints the
//                         value of FullLLKFunc::ComputeMixLLKs at each (%.17g)
//   --BenchEvals <n>      time n back-to-back ComputeMixLLKs calls at the start point
//   --BenchWarmup <w>     untimed calls before them (default 1)
//   --NoOptimize          skip OptimizeLLK (with --EvalPoints / --BenchEvals)
// and one machine-readable "VB2REF {json}" line on stdout per action.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "ContaminationEstimator.h"
#include "Error.h"

long g_vb2_ref_evals = 0;  // bumped by oracle/ref_hooks.h once per ComputeMixLLKs call

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Args {
  std::string ud = "Empty", mean = "Empty", bed = "Empty", svd = "Empty", pileup = "Empty";
  std::string out = "result", fixPC = "Empty", knownAF = "Empty", evalPoints = "Empty";
  double fixAlpha = -1., epsilon = 1e-8;
  bool within = false, disableSanity = false, verbose = false, noOptimize = false;
  int nPC = 2, nthread = 4, seed = 12345, benchEvals = 0, benchWarmup = 1;
};

bool parse(int argc, char **argv, Args &a) {
  for (int i = 1; i < argc; ++i) {
    std::string f = argv[i];
    auto val = [&](const char *name) -> const char * {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", name); exit(2); }
      return argv[++i];
    };
    if (f == "--SVDPrefix") a.svd = val("--SVDPrefix");
    else if (f == "--UDPath") a.ud = val("--UDPath");
    else if (f == "--MeanPath") a.mean = val("--MeanPath");
    else if (f == "--BedPath") a.bed = val("--BedPath");
    else if (f == "--PileupFile") a.pileup = val("--PileupFile");
    else if (f == "--Reference") (void)val("--Reference");
    else if (f == "--Output") a.out = val("--Output");
    else if (f == "--NumPC") a.nPC = atoi(val("--NumPC"));
    else if (f == "--NumThread") a.nthread = atoi(val("--NumThread"));
    else if (f == "--Seed") a.seed = atoi(val("--Seed"));
    else if (f == "--Epsilon") a.epsilon = atof(val("--Epsilon"));
    else if (f == "--FixPC") a.fixPC = val("--FixPC");
    else if (f == "--FixAlpha") a.fixAlpha = atof(val("--FixAlpha"));
    else if (f == "--KnownAF") a.knownAF = val("--KnownAF");
    else if (f == "--WithinAncestry") a.within = true;
    else if (f == "--DisableSanityCheck") a.disableSanity = true;
    else if (f == "--Verbose") a.verbose = true;
    else if (f == "--EvalPoints") a.evalPoints = val("--EvalPoints");
    else if (f == "--BenchEvals") a.benchEvals = atoi(val("--BenchEvals"));
    else if (f == "--BenchWarmup") a.benchWarmup = atoi(val("--BenchWarmup"));
    else if (f == "--NoOptimize") a.noOptimize = true;
    else { fprintf(stderr, "unknown option %s\n", f.c_str()); return false; }
  }
  if (a.svd != "Empty") { a.ud = a.svd + ".UD"; a.mean = a.svd + ".mu"; a.bed = a.svd + ".bed"; }
  return a.ud != "Empty" && a.mean != "Empty" && a.bed != "Empty" && a.pileup != "Empty";
}

// Reads that survive the marker skip rules of ComputeMixLLKs (ContaminationEstimator.h:238-249).
void used_counts(ContaminationEstimator &E, long &markers, long &reads) {
  markers = reads = 0;
  for (size_t i = 0; i < E.NumMarker; ++i) {
    const auto &rm = E.resolvedMarkers[i];
    if (rm.baseInfoIndex < 0) continue;
    size_t d = E.viewer.baseInfo[rm.baseInfoIndex].size();
    if (d == 0) continue;
    if (!E.isSanityCheckDisabled &&
        (d < (E.viewer.avgDepth - 3 * E.viewer.sdDepth) || d > (E.viewer.avgDepth + 3 * E.viewer.sdDepth)))
      continue;
    ++markers; reads += (long)d;
  }
}

void print_vec(const char *key, const std::vector<double> &v) {
  printf("\"%s\":[", key);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.17g", i ? "," : "", v[i]);
  printf("]");
}

}  // namespace

int main(int argc, char **argv) {
  Args a;
  if (!parse(argc, argv, a)) {
    fprintf(stderr, "usage: vb2_ref --SVDPrefix P --PileupFile F [reference CLI model flags] "
                    "[--EvalPoints file] [--BenchEvals n] [--NoOptimize]\n");
    return 2;
  }
  try {
    // main.cpp:283-319
    ContaminationEstimator Estimator(a.nPC, a.bed.c_str(), a.nthread, a.epsilon);
    Estimator.verbose = a.verbose;
    Estimator.seed = a.seed;
    Estimator.isHeter = !a.within;
    Estimator.isSanityCheckDisabled = a.disableSanity;
    if (a.fixPC != "Empty") {
      std::stringstream ss(a.fixPC);
      std::string token;
      std::vector<PCtype> tmpPC;
      while (std::getline(ss, token, ':')) tmpPC.push_back(atof(token.c_str()));
      if ((int)tmpPC.size() < a.nPC) error("parameter --fixPC provided smaller dimension than parameter --numPC(default value 2)");
      for (int i = 0; i < a.nPC; ++i) Estimator.PC[1][i] = tmpPC[i];
      Estimator.isPCFixed = true;
    } else if (fabs(a.fixAlpha + 1.) > std::numeric_limits<double>::epsilon()) {
      Estimator.alpha = a.fixAlpha;
      Estimator.isAlphaFixed = true;
    }
    if (a.knownAF != "Empty") {
      Estimator.isAFknown = true;
      Estimator.isPCFixed = true;
      Estimator.isHeter = false;
      Estimator.ReadAF(a.knownAF);
    }
    // main.cpp:321-333
    double t0 = now_s();
    Estimator.ReadSVDMatrix(a.ud, a.ud.substr(0, a.ud.size() - 3) + ".V", a.mean);
    double t1 = now_s();
    Estimator.ReadPileup(a.pileup);
    double t2 = now_s();
    // main.cpp:371-379
    if (!a.disableSanity) {
      if (Estimator.IsSanityCheckOK()) notice("Passing Marker Sanity Check...");
      else { warning("Insufficient Available markers"); return 1; }
    }
    printf("VB2REF {\"phase\":\"load\",\"panel_s\":%.6f,\"pileup_s\":%.6f,\"num_marker\":%u,"
           "\"avg_depth\":%.17g,\"sd_depth\":%.17g}\n",
           t1 - t0, t2 - t1, Estimator.NumMarker, Estimator.viewer.avgDepth, Estimator.viewer.sdDepth);

    if (a.evalPoints != "Empty" || a.benchEvals > 0) {
      Estimator.BuildResolvedMarkers();
      long mu = 0, ru = 0;
      used_counts(Estimator, mu, ru);
      if (a.evalPoints != "Empty") {
        std::ifstream fin(a.evalPoints);
        std::string line;
        while (std::getline(fin, line)) {
          if (line.empty() || line[0] == '#') continue;
          std::stringstream ss(line);
          std::vector<double> p1(a.nPC), p2(a.nPC);
          double alpha = 0;
          for (auto &x : p1) ss >> x;
          for (auto &x : p2) ss >> x;
          ss >> alpha;
          double v = Estimator.fn.ComputeMixLLKs(p1, p2, alpha);
          printf("VB2REF {\"phase\":\"eval\",\"llk\":%.17g}\n", v);
        }
      }
      if (a.benchEvals > 0) {
        std::vector<double> p1(a.nPC, 0.01), p2(a.nPC, 0.01);
        double sink = 0;
        for (int i = 0; i < a.benchWarmup; ++i) sink += Estimator.fn.ComputeMixLLKs(p1, p2, 0.03);  // warm-up
        double b0 = now_s();
        for (int i = 0; i < a.benchEvals; ++i) {
          p1[0] = 0.01 + 1e-6 * i;
          sink += Estimator.fn.ComputeMixLLKs(p1, p2, 0.03);
        }
        double b1 = now_s();
        printf("VB2REF {\"phase\":\"bench\",\"evals\":%d,\"seconds\":%.9f,\"threads\":%d,"
               "\"markers_used\":%ld,\"reads_used\":%ld,\"sink\":%.17g}\n",
               a.benchEvals, b1 - b0, a.nthread, mu, ru, sink);
      }
    }

    if (!a.noOptimize) {
      g_vb2_ref_evals = 0;
      double o0 = now_s();
      Estimator.OptimizeLLK(a.out);  // main.cpp:381-384
      double o1 = now_s();
      long mu = 0, ru = 0;
      used_counts(Estimator, mu, ru);
      // main.cpp:386-411 (.selfSM)
      {
        const char *headers = "#SEQ_ID\tRG\tCHIP_ID\t#SNPS\t#READS\tAVG_DP\tFREEMIX\tFREELK1\tFREELK0\tFREE_RH\tFREE_RA\tCHIPMIX\tCHIPLK1\tCHIPLK0\tCHIP_RH\tCHIP_RA\tDPREF\tRDPHET\tRDPALT";
        std::ofstream fout(a.out + ".selfSM");
        fout << headers << std::endl;
        fout << Estimator.viewer.SEQ_SM << "\tNA\tNA\t" << Estimator.NumMarker << "\t";
        if (Estimator.isPileupInput) fout << "NA";
        else fout << Estimator.viewer.numBases;
        fout << "\t" << Estimator.viewer.avgDepth << "\t"
             << ((Estimator.fn.globalAlpha < 0.5) ? Estimator.fn.globalAlpha : (1.f - Estimator.fn.globalAlpha)) << "\t"
             << -Estimator.fn.llk1 << "\t" << -Estimator.fn.llk0 << "\t" << "NA\tNA\t"
             << "NA\tNA\tNA\tNA\tNA\t" << "NA\tNA\tNA" << std::endl;
      }
      printf("VB2REF {\"phase\":\"optimize\",\"alpha\":%.17g,\"llk1\":%.17g,\"llk0\":%.17g,",
             Estimator.fn.globalAlpha, Estimator.fn.llk1, Estimator.fn.llk0);
      print_vec("pc_contam", Estimator.fn.globalPC); printf(",");
      print_vec("pc_intended", Estimator.fn.globalPC2);
      printf(",\"evals\":%ld,\"optimize_s\":%.6f,\"threads\":%d,\"markers_used\":%ld,\"reads_used\":%ld}\n",
             g_vb2_ref_evals, o1 - o0, a.nthread, mu, ru);
    }
  } catch (std::exception &e) {
    std::cerr << "Exiting due to ERROR:\n\t" << e.what() << std::endl;
    return -1;
  }
  return 0;
}
