"""The engine's C++ host side (libvb2host.so: readers, text-pileup parser, sanity filter, marker resolution,
Nelder-Mead) against the oracle's restatement of the reference -- no GPU needed."""
import numpy as np
import pytest

from verifybamid_b200 import host, panels, synth
from helpers import GOLD, HAPMAP, LONGREAD_PILEUP, RESULT_PILEUP, vo
import ctypes


def _same_problem(a, b):
    for f in ("ud", "means", "base_info_index", "info_offset", "bases", "quals"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    has = a.base_info_index >= 0
    assert np.array_equal(a.alt_base[has], b.alt_base[has])
    assert a.avg_depth == b.avg_depth and a.sd_depth == b.sd_depth and a.sanity_disabled == b.sanity_disabled


@pytest.mark.parametrize("pileup", [RESULT_PILEUP, LONGREAD_PILEUP])
def test_loader_matches_oracle_on_reference_fixtures(pileup):
    mine, summ = host.load_problem(HAPMAP, pileup, 2, disable_sanity=True)
    ref = vo.problem_from_files(HAPMAP, pileup, 2, disable_sanity=True)
    _same_problem(mine, ref)
    assert summ["num_marker"] == 9787


def test_parse_rules_and_quirks(tmp_path):
    """Indels, '^' read starts, '*'/'#' placeholders, '$', lines outside the panel, duplicate lines, depth-0
    lines, a short line that inherits fields (SimplePileupViewer.cpp:711-833)."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    sub = panels.PanelData("t", panel.ud[:40], panel.mu[:40], panel.chrom[:40], panel.pos[:40], panel.ref[:40],
                           panel.alt[:40], panel.v)
    prefix = panels.write_text_panel(sub, str(tmp_path / "p"))
    c, pos = sub.chrom, sub.pos
    lines = [
        "%s\t%d\tA\t9\t^].,+2ACg-1t*A$#Nn\tABCDEFGHI" % (c[0], pos[0]),
        "%s\t%d\tC\t3\t...\tIII" % (c[1], pos[1]),
        "%s\t%d\tC\t2\tAA\tII" % (c[1], pos[1]),                 # duplicate line: counted, data discarded
        "%s\t%d\tG\t0\t*\t*" % (c[2], pos[2]),                   # depth-0 column
        "%s\t%d\tG\t4\tACGT\tIIII" % (c[3], pos[3] + 1),         # not a panel position
        "Z\t5\tG\t4\tACGT\tIIII",                                # not a panel chromosome
        "%s\t%d\tT\t5\tacgtn\t!~5{I" % (c[5], pos[5]),           # qualities at both ends of the range
        "%s\t%d\tT\t2\t<>\tII" % (c[6], pos[6]),                 # reference skips: ignored, no quality consumed
    ]
    pile = tmp_path / "x.pileup"
    pile.write_text("\n".join(lines) + "\n")
    mine, summ = host.load_problem(prefix, str(pile), 2, disable_sanity=True)
    ref = vo.problem_from_files(prefix, str(pile), 2, disable_sanity=True)
    _same_problem(mine, ref)
    assert summ["effective_num_site"] == 6 and summ["num_bases"] == 6 + 3 + 2 + 0 + 5 + 0
    assert bytes(mine.bases[:6]) == b".,gANn" and bytes(mine.quals[:6]) == b"ABCEGH"
    d = mine.depths()
    assert d[0] == 6 and d[1] == 3 and d[2] == 0 and d[5] == 5 and d[6] == 0


def test_sanity_filter_and_known_af(tmp_path):
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=15.0, alpha=0.02, seed=21, n_markers=3000)
    prefix = panels.write_text_panel(s.panel, str(tmp_path / "p"))
    pile = s.write_pileup(str(tmp_path / "s.pileup"))
    mine, summ = host.load_problem(prefix, pile, 2, disable_sanity=False)
    ref = vo.problem_from_files(prefix, pile, 2, disable_sanity=False)
    _same_problem(mine, ref)
    assert summ["sanity_ok"] and mine.sd_depth > 0
    # the generator's direct arrays are what the parser yields
    assert np.array_equal(mine.bases, s.problem.bases) and np.array_equal(mine.quals, s.problem.quals)
    assert mine.avg_depth == s.problem.avg_depth and abs(mine.sd_depth - s.problem.sd_depth) < 1e-12
    assert mine.used_counts() == s.problem.used_counts() == ref.used_counts()
    # --KnownAF file: chr x pos ref alt AF (ContaminationEstimator.cpp:461-487)
    kaf = tmp_path / "af.txt"
    kaf.write_text("".join("%s\t%d\t%d\t%s\t%s\t%r\n" % (c, p - 1, p, r, a, 0.01 * (i % 90))
                           for i, (c, p, r, a) in enumerate(zip(s.panel.chrom, s.panel.pos, s.panel.ref, s.panel.alt))))
    mine, _ = host.load_problem(prefix, pile, 2, disable_sanity=True, known_af=str(kaf))
    ref = vo.problem_from_files(prefix, pile, 2, disable_sanity=True, known_af_path=str(kaf))
    assert np.array_equal(mine.known_af, ref.known_af)


def _oracle_amoeba(fn, start, ftol):
    lib = vo.lib()
    CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int)
    lib.vb2o_amoeba_run.restype = ctypes.c_double
    lib.vb2o_amoeba_run.argtypes = [CB, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                    ctypes.POINTER(ctypes.c_long)]
    pt = np.ascontiguousarray(start, dtype=np.float64).copy()
    cyc = ctypes.c_long()
    cb = CB(lambda u, v, n: float(fn(np.ctypeslib.as_array(v, shape=(n,)).copy())))
    r = lib.vb2o_amoeba_run(cb, None, pt.size, pt.ctypes.data, ftol, ctypes.byref(cyc))
    return float(r), pt, int(cyc.value)


@pytest.mark.parametrize("dim", [1, 3, 5])
def test_nelder_mead_visits_the_reference_points(dim):
    """Same function values in -> same simplex trajectory out as the restated AmoebaMinimizer, bit for bit."""
    def rosen(v):
        return float(sum(100.0 * (v[i + 1] - v[i] ** 2) ** 2 + (1 - v[i]) ** 2 for i in range(len(v) - 1)) + (v[0] - 0.3) ** 2)
    seen_a, seen_b = [], []
    ra = host.amoeba_minimize(lambda v: (seen_a.append(v.tolist()), rosen(v))[1], [0.01] * dim, 1e-8)
    rb = _oracle_amoeba(lambda v: (seen_b.append(v.tolist()), rosen(v))[1], [0.01] * dim, 1e-8)
    assert seen_a == seen_b and len(seen_a) > dim + 1
    assert ra[0] == rb[0] and ra[1].tolist() == rb[1].tolist() and ra[2] == rb[2]


def test_nelder_mead_gives_up_after_cycle_max():
    calls = [0]
    def noisy(v):                       # never converges: relative spread stays large
        calls[0] += 1
        return float((-1) ** calls[0] * (1 + calls[0] % 7))
    r, _, cyc = host.amoeba_minimize(noisy, [0.0, 0.0], 1e-8)
    assert r == np.finfo(np.float64).max and cyc > 50000      # MathGenMin.cpp:380-383


def test_cohort_lock_step_keeps_every_trajectory():
    """Cohort mode: samples optimise concurrently, the coordinator serves all pending requests with one launch;
    each sample must land exactly where it lands alone, and the launch count is the longest sample's eval count."""
    def rosen(v):
        return float(sum(100.0 * (v[i + 1] - v[i] ** 2) ** 2 + (1 - v[i]) ** 2 for i in range(len(v) - 1)) + (v[0] - 0.3) ** 2)
    n, dim = 7, 3
    starts = np.array([[0.01 + 0.02 * i] * dim for i in range(n)])
    launches, pts, fmin, cyc = host.cohort_selftest(rosen, starts, 1e-8)
    evals = []
    for i in range(n):
        count = [0]
        def f_i(v, i=i):
            count[0] += 1
            return rosen(v - 0.1 * i)
        r, p, c = host.amoeba_minimize(f_i, starts[i], 1e-8)
        assert r == fmin[i] and p.tolist() == pts[i].tolist() and c == cyc[i]
        evals.append(count[0])
    assert launches == max(evals)            # one launch per lock-step round, samples drop out as they converge
    assert len(set(evals)) > 1               # (the samples really need different numbers of steps)
