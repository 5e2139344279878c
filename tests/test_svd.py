"""Panel construction: SVDcalculator::ComputeSvdGram and the centring in front of it
(reference SVDcalculator.cpp:258-339, :402-409) on the device, through the C ABI of libvb2svd.so.

The checker is the reference's OWN code -- oracle/_ref/vb2_svd_ref = SVDcalculator.cpp + libVcf + Eigen compiled
unmodified -- and oracle/svd_oracle.py (numpy restatement, pinned against that binary here).  The comparison is the
reference test's (TestGramSVD.cpp:55-73): per column, sign-aligned, max deviation relative to max(column norm, norm of
the dominant column).  Tolerance: the reference allows 1e-2 between its two fp32 decompositions and observes ~1e-5;
here GRAM_TOL = 1e-3 (measured 1e-6 .. 1e-5: fp32 accumulation order and a different eigensolver)."""
import ctypes

import numpy as np
import pytest

from verifybamid_b200 import svd
from oracle import svd_oracle as so

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
