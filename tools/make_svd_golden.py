#!/usr/bin/env python
"""Golden vectors for the panel-construction parity tests, generated HERE from the reference's own code
(oracle/_ref/vb2_svd_ref = SVDcalculator.cpp + libVcf + Eigen, unmodified) -- the script that made
tests/golden/svd/gram_300x40.npz and tests/golden/svd/vcf_panel.npz.   python tools/make_svd_golden.py"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svd_oracle as so

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "svd")
os.makedirs(out, exist_ok=True)
assert so.reference_available(), "build oracle/_ref first (make -C oracle ref)"

# 1. ComputeSvdGram / ComputeSvdJacobi on the LCG test matrix (TestGramSVD.cpp's generator), 300 x 40, 6 components
g = so.lcg_genotypes(300, 40, 2024)
a, mu = so.center(g)
ud, pc, sv = so.reference_svd(a, 6, "gram")
jud, jpc, jsv = so.reference_svd(a, 6, "jacobi")
np.savez_compressed(os.path.join(out, "gram_300x40.npz"), genotype=g, mu=mu, gram_ud=ud, gram_pc=pc, gram_sv=sv,
                    jacobi_ud=jud, jacobi_pc=jpc, jacobi_sv=jsv)

# 2. ReadVcf + ProcessRefVCF on the synthetic panel VCF (oracle/svd_oracle.py:write_test_vcf, seed 1): the genotype
#    matrix's checksum, the first rows of .mu/.bed and the .UD/.V tables
with tempfile.TemporaryDirectory() as td:
    vcf = os.path.join(td, "panel.vcf")
    so.write_test_vcf(vcf)
    chrs = [str(i) for i in range(1, 23)]
    geno = so.reference_read_vcf(vcf, chrs)
    so.reference_process_vcf(vcf, 10, True, True, chrs)
    mu_lines = open(vcf + ".mu").read().splitlines()
    bed_lines = open(vcf + ".bed").read().splitlines()
    ud = np.array([[float(x) for x in l.rstrip("\t\n").split("\t")] for l in open(vcf + ".UD")])
    v = np.array([[float(x) for x in l.rstrip("\t\n").split("\t")[1:]] for l in open(vcf + ".V")])
np.savez_compressed(os.path.join(out, "vcf_panel.npz"), genotype=geno, mu_text=np.array(mu_lines), bed_text=np.array(bed_lines),
                    ud=ud.astype(np.float32), v=v.astype(np.float32))
print("written:", os.listdir(out), [os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)])
