// oracle/svd_ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Command-line front to the reference's OWN panel-construction code (SVDcalculator.{h,cpp}, compiled unmodified from
// /root/reference together with its libVcf and Eigen): the checker for verifybamid_b200/csrc/svd_gram.cu and svd_panel.cpp.
//   vb2_svd_ref gram|jacobi M N K in.f32 out.f32
//        in : the mean-centred matrix, row-major [M][N] fp32   (SVDcalculator::ComputeSvdGram / ComputeSvdJacobi,
//        out: UD [M][K], PC [N][K], singular values [N or min(M,N)] fp32, row-major    SVDcalculator.cpp:258-361)
//   vb2_svd_ref vcf <ref.vcf> <numSVDPCs> <gram 0|1> <skipMinSampleCountCheck 0|1> [includeChr,comma,separated]
//        SVDcalculator::ProcessRefVCF (cpp:363-449): writes <ref.vcf>.UD/.mu/.bed/.V exactly as `--RefVCF` does
//   vb2_svd_ref readvcf <ref.vcf> out.bin [includeChr]
//        SVDcalculator::ReadVcf (cpp:22-224): int32 nMarkers, int32 nSamples, then the matrix [nMarkers][nSamples] int8
#include "SVDcalculator.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

using Eigen::MatrixXf;
using Eigen::VectorXf;

static int die(const char *msg) {
  fprintf(stderr, "vb2_svd_ref: %s\n", msg);
  return 2;
}

int main(int argc, char **argv) {
  if (argc >= 7 && (!strcmp(argv[1], "gram") || !strcmp(argv[1], "jacobi"))) {
    const int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]);
    std::vector<float> in((size_t)M * N);
    FILE *f = fopen(argv[5], "rb");
    if (!f || fread(in.data(), sizeof(float), in.size(), f) != in.size()) return die("cannot read the input matrix");
    fclose(f);
    MatrixXf A(M, N);
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) A(i, j) = in[(size_t)i * N + j];
    MatrixXf ud, pc;
    VectorXf sv;
    if (!strcmp(argv[1], "gram")) SVDcalculator::ComputeSvdGram(A, K, ud, pc, sv);
    else SVDcalculator::ComputeSvdJacobi(A, K, ud, pc, sv);
    f = fopen(argv[6], "wb");
    if (!f) return die("cannot write the output");
    for (int i = 0; i < M; ++i)
      for (int c = 0; c < K; ++c) { float v = ud(i, c); fwrite(&v, 4, 1, f); }
    for (int j = 0; j < N; ++j)
      for (int c = 0; c < K; ++c) { float v = pc(j, c); fwrite(&v, 4, 1, f); }
    const int ns = (int)sv.size();
    fwrite(&ns, 4, 1, f);
    for (int i = 0; i < ns; ++i) { float v = sv(i); fwrite(&v, 4, 1, f); }
    fclose(f);
    return 0;
  }
  if (argc >= 6 && !strcmp(argv[1], "vcf")) {
    std::unordered_set<std::string> chr;
    if (argc >= 7) {
      std::stringstream ss(argv[6]);
      std::string tok;
      while (std::getline(ss, tok, ','))
        if (!tok.empty()) chr.insert(tok);
    }
    SVDcalculator calc;
    calc.ProcessRefVCF(argv[2], chr, atoi(argv[5]) != 0, atoi(argv[3]), atoi(argv[4]) != 0);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "readvcf")) {  // SVDcalculator::ReadVcf alone: the genotype matrix it builds
    std::unordered_set<std::string> chr;
    if (argc >= 5) {
      std::stringstream ss(argv[4]);
      std::string tok;
      while (std::getline(ss, tok, ','))
        if (!tok.empty()) chr.insert(tok);
    }
    SVDcalculator calc;
    std::vector<std::vector<char> > genotype;
    int nS = 0, nM = 0;
    calc.ReadVcf(argv[2], genotype, nS, nM, chr);
    FILE *f = fopen(argv[3], "wb");
    if (!f) return die("cannot write the output");
    fwrite(&nM, 4, 1, f);
    fwrite(&nS, 4, 1, f);
    for (int i = 0; i < nM; ++i) fwrite(genotype[i].data(), 1, (size_t)nS, f);
    fclose(f);
    return 0;
  }
  return die("usage: gram|jacobi M N K in out  |  vcf path numSVDPCs gram skipCheck [includeChr]  |  readvcf path out [includeChr]");
}
