"""oracle/svd_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
CPU restatement (numpy, single precision) of the reference's panel construction:
  center()           ProcessRefVCF's centring                      reference SVDcalculator.cpp:402-409
  compute_svd_gram() SVDcalculator::ComputeSvdGram                  reference SVDcalculator.cpp:258-339
  lcg_genotypes()    the deterministic test matrix of TestGramSVD   reference TestGramSVD.cpp:32-41 (LCG), makeGenotypeMatrix
  column_error()     the sign-aligned, scale-aware column comparison of TestGramSVD.cpp:55-73
Pinned against the reference's own code (oracle/_ref/vb2_svd_ref) in tests/test_svd.py; only tests may import it."""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "vb2_svd_ref")


def center(genotype):
    g = np.asarray(genotype, dtype=np.float32)
    mu = (g.sum(axis=1, dtype=np.float32) / np.float32(g.shape[1])).astype(np.float32)   # rowwise().mean()
    return (g - mu[:, None]).astype(np.float32), mu


def compute_svd_gram(centered, n_pc):
    a = np.asarray(centered, dtype=np.float32)
    gram = (a.T @ a).astype(np.float32)                      # G = A^T A            (cpp:305-306)
    w, v = np.linalg.eigh(gram)                              # ascending            (cpp:309)
    sv = np.sqrt(np.maximum(w[::-1], 0)).astype(np.float32)  # descending, clamped  (cpp:322)
    pc = v[:, ::-1][:, :n_pc].astype(np.float32)             # top eigenvectors     (cpp:337)
    return (a @ pc).astype(np.float32), pc, sv               # UD = A * PC          (cpp:338)


def lcg_genotypes(m, n, seed):
    """Genotypes in {0,1,2} from the reference test's LCG (state*1103515245+12345, bits 16..30), with a per-marker allele
    frequency so that the matrix has structure."""
    state = np.uint32(seed)
    out = np.zeros((m, n), np.int8)

    def nxt():
        nonlocal state
        state = np.uint32((int(state) * 1103515245 + 12345) & 0xFFFFFFFF)
        return (int(state) >> 16) & 0x7FFF
    for i in range(m):
        af = 0.05 + 0.9 * (nxt() / 32768.0)
        for j in range(n):
            out[i, j] = (nxt() / 32768.0 < af) + (nxt() / 32768.0 < af)
    return out


def column_error(a, b, ref_scale):
    """TestGramSVD.cpp:55-73: max |a - sign*b| / max(||a||, ref_scale), sign from the dot product."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    sign = 1.0 if float(a @ b) >= 0 else -1.0
    denom = max(float(np.linalg.norm(a)), float(ref_scale))
    return float(np.abs(a - sign * b).max()) / denom if denom > 0 else float(np.abs(a - sign * b).max())


def reference_available():
    return os.access(REF_BIN, os.X_OK)


def reference_svd(centered, n_pc, method="gram"):
    """The reference's own ComputeSvdGram / ComputeSvdJacobi on `centered` (through oracle/_ref/vb2_svd_ref)."""
    a = np.ascontiguousarray(centered, dtype=np.float32)
    m, n = a.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.f32"), os.path.join(td, "out.f32")
        a.tofile(fin)
        subprocess.run([REF_BIN, method, str(m), str(n), str(n_pc), fin, fout], check=True, capture_output=True)
        raw = np.fromfile(fout, dtype=np.float32)
    ud = raw[:m * n_pc].reshape(m, n_pc)
    pc = raw[m * n_pc:(m + n) * n_pc].reshape(n, n_pc)
    ns = int(raw[(m + n) * n_pc:(m + n) * n_pc + 1].view(np.int32)[0])
    sv = raw[(m + n) * n_pc + 1:(m + n) * n_pc + 1 + ns]
    return ud, pc, sv
