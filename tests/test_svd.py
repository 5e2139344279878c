"""Panel construction: SVDcalculator::ComputeSvdGram and the centring in front of it
(reference SVDcalculator.cpp:258-339, :402-409) on the device, through the C ABI of libvb2svd.so.

The checker is the reference's OWN code -- oracle/_ref/vb2_svd_ref = SVDcalculator.cpp + libVcf + Eigen compiled
unmodified -- and oracle/svd_oracle.py (numpy restatement, pinned against that binary here).  The comparison is the
reference test's (TestGramSVD.cpp:55-73): per column, sign-aligned, max deviation relative to max(column norm, norm of
the dominant column).  Tolerance: the reference allows 1e-2 between its two fp32 decompositions and observes ~1e-5;
here GRAM_TOL = 1e-3 (measured 1e-6 .. 1e-5: fp32 accumulation order and a different eigensolver)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from verifybamid_b200 import host, svd
from oracle import svd_oracle as so
from helpers import ROOT

AUTOSOMES = [str(i) for i in range(1, 23)]

GRAM_TOL = 1e-3
needs_ref = pytest.mark.skipif(not so.reference_available(), reason="oracle/_ref/vb2_svd_ref not built")


def structured_genotypes(m, n, seed, n_pop=3):
    """Genotypes with population structure (a few clearly separated top singular values) and ~1 % missing (-1)."""
    rng = np.random.default_rng(seed)
    pop = rng.integers(0, n_pop, n)
    af = np.clip(rng.uniform(0.05, 0.95, (m, 1)) + rng.normal(0, 0.15, (m, n_pop)), 0.01, 0.99)
    p = af[:, pop]
    g = (rng.random((m, n)) < p).astype(np.int8) + (rng.random((m, n)) < p).astype(np.int8)
    g[rng.random((m, n)) < 0.01] = -1
    return g


def compare(got_ud, got_pc, got_sv, ref_ud, ref_pc, ref_sv, k, tol):
    scale = float(np.linalg.norm(ref_ud[:, 0]))
    for c in range(k):
        assert so.column_error(ref_ud[:, c], got_ud[:, c], scale) <= tol, ("UD", c)
        assert so.column_error(ref_pc[:, c], got_pc[:, c], 1.0) <= tol, ("PC", c)
    # the top k singular values (the trailing ones of a Gram matrix are ill-determined: sqrt of rounding noise, cpp:285-290)
    assert np.abs(got_sv[:k] - ref_sv[:k]).max() <= tol * float(ref_sv[0])


@needs_ref
@pytest.mark.parametrize("m,n,k", [(300, 40, 5), (1200, 96, 10)])
def test_numpy_restatement_matches_the_reference_code(m, n, k):
    a, _ = so.center(structured_genotypes(m, n, seed=m + n))
    ud, pc, sv = so.compute_svd_gram(a, k)
    rud, rpc, rsv = so.reference_svd(a, k, "gram")
    compare(ud, pc, sv, rud, rpc, rsv, k, 1e-4)
    jud, jpc, jsv = so.reference_svd(a, k, "jacobi")      # the reference's two decompositions agree (its own test)
    compare(rud, rpc, rsv[:len(jsv)], jud, jpc, jsv, k, 1e-2)


def test_lcg_matrix_is_the_reference_tests_generator():
    g = so.lcg_genotypes(6, 5, 42)
    assert g.shape == (6, 5) and set(np.unique(g)) <= {0, 1, 2}
    assert (g == so.lcg_genotypes(6, 5, 42)).all() and (g != so.lcg_genotypes(6, 5, 43)).any()


def test_library_loads_and_refuses_bad_arguments_without_touching_a_device():
    lib = svd.load_library()
    assert hasattr(lib, "vb2_svd_gram") and hasattr(lib, "vb2_svd_last_error")
    assert ctypes.sizeof(svd._Desc) == 80
    d = svd._Desc()
    d.struct_size = 8
    assert lib.vb2_svd_gram(ctypes.byref(d)) == 1 and b"struct_size" in lib.vb2_svd_last_error()
    g = np.zeros((4, 3), np.int8)
    with pytest.raises(svd.SVDError) as ei:       # numPCs out of [1, min(M, N)]  (cpp:299-302)
        svd.svd_gram(g, 4)
    assert ei.value.code == 1 and "numPCs" in str(ei.value)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("m,n,k", [(300, 40, 5), (5000, 200, 10), (2111, 130, 10), (1500, 257, 40)])
def test_device_gram_svd_matches_the_reference_code(m, n, k):
    g = structured_genotypes(m, n, seed=3 * m + n)
    r = svd.svd_gram(g, k)
    a, mu = so.center(g)
    assert (r["mu"] == mu).all()                                  # integer sums: the mean is exact, bit for bit
    rud, rpc, rsv = so.reference_svd(a, k, "gram")
    # the structured components to GRAM_TOL; beyond them the eigenvalues of the noise bulk are nearly degenerate and two
    # fp32 eigensolvers pick slightly different directions: the reference test's own bound (1e-2, TestGramSVD.cpp:19-21)
    compare(r["ud"], r["pc"], r["singular"], rud, rpc, rsv, min(k, 10), GRAM_TOL)
    compare(r["ud"], r["pc"], r["singular"], rud, rpc, rsv, k, 1e-2)
    # ComputeSvdGram's own argument (an already centred matrix) gives the same decomposition
    r2 = svd.svd_gram(a, k)
    compare(r2["ud"], r2["pc"], r2["singular"], rud, rpc, rsv, min(k, 10), GRAM_TOL)
    # UD . PC products are what the likelihood uses: independent of the eigensolver's sign choice
    top = min(k, 3)
    assert np.abs(r["ud"][:50, :top] @ r["pc"][:, :top].T - rud[:50, :top] @ rpc[:, :top].T).max() <= GRAM_TOL * float(rsv[0])


@pytest.mark.gpu
def test_device_gram_svd_small_shapes_and_reproducibility():
    g = so.lcg_genotypes(6, 4, 11)
    r = svd.svd_gram(g, 4)                                         # numPCs = min(M, N)
    a, mu = so.center(g)
    ud, pc, sv = so.compute_svd_gram(a, 4)
    assert (r["mu"] == mu).all() and np.abs(r["singular"] - sv).max() <= 1e-4 * sv[0]
    for c in range(3):                                             # (the last component of a centred matrix is ~0)
        assert so.column_error(ud[:, c], r["ud"][:, c], float(np.linalg.norm(ud[:, 0]))) <= GRAM_TOL
    g = structured_genotypes(4000, 100, seed=5)
    r1, r2 = svd.svd_gram(g, 8), svd.svd_gram(g, 8)                # no atomics: the same bits every time
    assert (r1["ud"] == r2["ud"]).all() and (r1["singular"] == r2["singular"]).all()


@pytest.fixture(scope="module")
def test_vcf(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("refvcf") / "panel.vcf")
    so.write_test_vcf(path)
    return path


@needs_ref
def test_host_vcf_reader_builds_the_reference_genotype_matrix(test_vcf):
    """ReadVcf (cpp:22-224) of csrc/svd_panel.cpp against the reference's own: FILTER / multi-allelic / non-SNP /
    chromosome / missing-rate rules, PL > GL > GT priority, unparsed samples kept as -1 -- the same matrix, bit for bit."""
    for chrs in (AUTOSOMES, []):
        want = so.reference_read_vcf(test_vcf, chrs)
        got = host.read_vcf(test_vcf, chrs)
        assert got["genotype"].shape == want.shape and (got["genotype"] == want).all()
        assert want.shape[1] == 64 and 5000 < want.shape[0] < 5400 and (want == -1).any()
    assert set(got["chrom"]) == {"1", "2", "X"} and len(got["pos"]) == want.shape[0]
    assert set(host.read_vcf(test_vcf, AUTOSOMES)["chrom"]) == {"1", "2"}
    import gzip, shutil                                           # a compressed panel reads the same (zlib)
    with open(test_vcf, "rb") as fi, gzip.open(test_vcf + ".gz", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    assert (host.read_vcf(test_vcf + ".gz", AUTOSOMES)["genotype"] == so.reference_read_vcf(test_vcf, AUTOSOMES)).all()


def test_host_vcf_reader_reports_malformed_input(tmp_path, capfd):
    bad = tmp_path / "bad.vcf"
    bad.write_text("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\tB\n"
                   "1\t100\t.\tA\tC\t.\tPASS\t.\tGT:PL\t0/0\t0/1:3,0,9\n")
    with pytest.raises(RuntimeError):                            # libVcfFile.cpp:931-934 (error() prints, then throws)
        host.read_vcf(str(bad))
    assert "do not match" in capfd.readouterr().err
    dup = tmp_path / "dup.vcf"
    dup.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tA\n"
                   "1\t100\t.\tA\tC\t.\tPASS\t.\tGT\t0/0\n1\t100\t.\tA\tG\t.\tPASS\t.\tGT\t0/1\n")
    with pytest.raises(RuntimeError):                            # cpp:60-63
        host.read_vcf(str(dup))
    assert "Duplicated Marker" in capfd.readouterr().err


def _read_table(path, skip_first_column=False):
    rows = [line.rstrip("\t\n").split("\t") for line in open(path)]
    return np.array([[float(x) for x in (r[1:] if skip_first_column else r)] for r in rows])


@pytest.mark.gpu
@needs_ref
def test_cli_refvcf_writes_the_reference_panel_files(test_vcf, tmp_path):
    """`--RefVCF` end to end (main.cpp:232-257, ProcessRefVCF cpp:363-449, WriteSVD cpp:471-513): .bed and .mu are the
    reference's byte for byte (the means are exact); .UD and .V agree column by column up to the eigensolver's sign."""
    import shutil
    ours, ref = str(tmp_path / "ours.vcf"), str(tmp_path / "ref.vcf")
    shutil.copy(test_vcf, ours); shutil.copy(test_vcf, ref)
    so.reference_process_vcf(ref, 10, True, True, AUTOSOMES)
    cli = os.path.join(ROOT, "verifybamid_b200", "VerifyBamID")
    r = subprocess.run([cli, "--RefVCF", ours, "--SkipMinSampleCountCheck", "--NumSVDPCs", "10", "--GramSVD"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Success!" in r.stderr and "eigendecomposition" in r.stderr
    assert open(ours + ".bed").read() == open(ref + ".bed").read()
    assert open(ours + ".mu").read() == open(ref + ".mu").read()
    ud, rud = _read_table(ours + ".UD"), _read_table(ref + ".UD")
    v, rv = _read_table(ours + ".V", True), _read_table(ref + ".V", True)
    assert ud.shape == rud.shape == (sum(1 for _ in open(ref + ".bed")), 10) and v.shape == rv.shape == (64, 10)
    assert [l.split("\t")[0] for l in open(ours + ".V")] == [l.split("\t")[0] for l in open(ref + ".V")]
    scale = float(np.linalg.norm(rud[:, 0]))
    for c in range(10):        # (files carry 6 significant digits)
        tol = GRAM_TOL if c < 3 else 1e-2
        assert so.column_error(rud[:, c], ud[:, c], scale) <= tol, ("UD", c)
        assert so.column_error(rv[:, c], v[:, c], 1.0) <= tol, ("V", c)
    # without --SkipMinSampleCountCheck a 64-sample panel is refused, as in the reference (cpp:383-396)
    r = subprocess.run([cli, "--RefVCF", ours], capture_output=True, text=True)
    assert r.returncode != 0 and "Insufficient number of individuals" in r.stderr


# ---- committed golden vectors (tests/golden/svd/*.npz, made from the reference's own code by tools/make_svd_golden.py) ----
GOLD = os.path.join(ROOT, "tests", "golden", "svd")


def test_restatement_and_host_reader_match_the_committed_goldens(tmp_path):
    z = np.load(os.path.join(GOLD, "gram_300x40.npz"))
    g = z["genotype"]
    assert (g == so.lcg_genotypes(300, 40, 2024)).all()
    a, mu = so.center(g)
    assert (mu == z["mu"]).all()
    ud, pc, sv = so.compute_svd_gram(a, 6)
    compare(ud, pc, sv, z["gram_ud"], z["gram_pc"], z["gram_sv"], 6, 1e-4)
    compare(z["gram_ud"], z["gram_pc"], z["gram_sv"], z["jacobi_ud"], z["jacobi_pc"], z["jacobi_sv"], 6, 1e-2)   # TestGramSVD's own claim
    v = np.load(os.path.join(GOLD, "vcf_panel.npz"))
    vcf = str(tmp_path / "panel.vcf")
    so.write_test_vcf(vcf)
    got = host.read_vcf(vcf, AUTOSOMES)
    assert got["genotype"].shape == v["genotype"].shape and (got["genotype"] == v["genotype"]).all()
    bed = ["%s\t%d\t%d\t%s\t%s" % (c, p - 1, p, r, a_) for c, p, r, a_ in zip(got["chrom"], got["pos"], got["ref"], got["alt"])]
    assert bed == list(v["bed_text"])


@pytest.mark.gpu
def test_device_gram_svd_and_cli_match_the_committed_goldens(tmp_path):
    z = np.load(os.path.join(GOLD, "gram_300x40.npz"))
    r = svd.svd_gram(z["genotype"], 6)
    assert (r["mu"] == z["mu"]).all()
    compare(r["ud"], r["pc"], r["singular"], z["gram_ud"], z["gram_pc"], z["gram_sv"], 6, GRAM_TOL)
    v = np.load(os.path.join(GOLD, "vcf_panel.npz"))
    vcf = str(tmp_path / "panel.vcf")
    so.write_test_vcf(vcf)
    cli = os.path.join(ROOT, "verifybamid_b200", "VerifyBamID")
    p = subprocess.run([cli, "--RefVCF", vcf, "--SkipMinSampleCountCheck"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    assert open(vcf + ".mu").read().splitlines() == list(v["mu_text"])
    assert open(vcf + ".bed").read().splitlines() == list(v["bed_text"])
    ud, vv = _read_table(vcf + ".UD"), _read_table(vcf + ".V", True)
    scale = float(np.linalg.norm(v["ud"][:, 0]))
    for c in range(10):
        tol = GRAM_TOL if c < 3 else 1e-2
        assert so.column_error(v["ud"][:, c], ud[:, c], scale) <= tol and so.column_error(v["v"][:, c], vv[:, c], 1.0) <= tol
