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


def reference_read_vcf(path, include_chr=()):
    """The reference's own SVDcalculator::ReadVcf: the genotype matrix [markers][samples] int8."""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "g.bin")
        subprocess.run([REF_BIN, "readvcf", path, out] + ([",".join(include_chr)] if include_chr else []), check=True,
                       capture_output=True)
        raw = np.fromfile(out, dtype=np.int8)
    nm, ns = (int(x) for x in raw[:8].view(np.int32))
    return raw[8:].reshape(nm, ns)


def reference_process_vcf(path, n_pcs=10, gram=True, skip_min_sample_check=True, include_chr=()):
    """The reference's own ProcessRefVCF (what `--RefVCF` runs): writes path.UD / .mu / .bed / .V."""
    subprocess.run([REF_BIN, "vcf", path, str(n_pcs), "1" if gram else "0", "1" if skip_min_sample_check else "0"] +
                   ([",".join(include_chr)] if include_chr else []), check=True, capture_output=True)


def write_test_vcf(path, n_marker=5400, n_sample=64, seed=1):
    """A plain-text reference-panel VCF that exercises every rule of ReadVcf (cpp:22-224): FILTER, multi-allelic and
    indel records, an excluded chromosome, GT / PL / GL in every priority combination, missing samples (kept as -1) and
    markers over the 20 % missing-rate limit.  Returns how many records were written."""
    rng = np.random.default_rng(seed)
    pop = rng.integers(0, 3, n_sample)
    bases = "ACGT"
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.1\n##source=vb2-test\n")
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("S%03d" % i for i in range(n_sample)) + "\n")
        for i in range(n_marker):
            chrom = "1" if i < n_marker // 2 else ("2" if i < n_marker - 60 else "X")
            pos = 1000 + 37 * i
            ref = bases[i % 4]
            alt = bases[(i + 1 + i // 7 % 3) % 4]
            flt = "PASS"
            kind = i % 97
            if kind == 5: flt = "LowQual"
            if kind == 6: flt = "PASS;q10"
            if kind == 7: alt = alt + "," + bases[(i + 2) % 4] if bases[(i + 2) % 4] != ref else alt + ",N"
            if kind == 8: ref = ref + "T"
            if kind == 9: alt = alt + "G"
            af = np.clip(rng.uniform(0.05, 0.95) + rng.normal(0, 0.2, 3), 0.02, 0.98)
            p = af[pop]
            g = (rng.random(n_sample) < p).astype(int) + (rng.random(n_sample) < p).astype(int)
            fmt_kind = i % 5      # 0,1: GT   2: GT:PL   3: GL:GT   4: GT:DP:PL with some PL missing
            fmt = ["GT", "GT", "GT:PL", "GL:GT", "GT:DP:PL"][fmt_kind]
            miss_rate = 0.5 if kind == 11 else (0.15 if kind == 12 else 0.01)
            vals = []
            for j in range(n_sample):
                gt = ["0/0", "0|1", "1/1"][g[j]] if (i + j) % 3 else ["0|0", "1/0", "1|1"][g[j]]
                if rng.random() < miss_rate:
                    vals.append("./." if rng.random() < 0.7 else ".")
                    continue
                pl = [[0, 33, 255], [40, 0, 45], [300, 28, 0]][g[j]]
                if fmt_kind < 2:
                    vals.append(gt)
                elif fmt_kind == 2:
                    vals.append("%s:%d,%d,%d" % (gt, *pl))
                elif fmt_kind == 3:
                    vals.append("%.2f,%.2f,%.2f:%s" % (-pl[0] / 10.0, -pl[1] / 10.0, -pl[2] / 10.0, gt))
                else:
                    vals.append("%s:%d:%s" % (gt, 7 + j % 5, ".,.,." if j % 11 == 0 else "%d,%d,%d" % tuple(pl)))
            f.write("%s\t%d\trs%d\t%s\t%s\t50\t%s\t.\t%s\t%s\n" % (chrom, pos, i, ref, alt, flt, fmt, "\t".join(vals)))
    return n_marker
