"""The CPU oracle (oracle/) against every golden vector the reference's tests hold for this path
 and, when present, against the reference's own sources compiled here
(oracle/_ref/vb2_ref)."""
import os

import numpy as np
import pytest

from helpers import (GOLD, HAPMAP, KAT_LONGREAD, KAT_POINTS, KAT_RESULT, LONGREAD_PILEUP, RESULT_PILEUP,
                     golden_problem, vo)

EXPECTED = os.path.join(GOLD, "expected")


@pytest.mark.parametrize("pileup,kat", [(RESULT_PILEUP, KAT_RESULT), (LONGREAD_PILEUP, KAT_LONGREAD)])
def test_known_answer_llk(pileup, kat):
    p = golden_problem(pileup)
    for (pc1, pc2, a), want in zip(KAT_POINTS, kat):
        got = p.compute_mix_llks(pc1, pc2, a)
        assert abs(got - want) <= 1e-12 * abs(want), (got, want)


def test_fixture_facts():
    # 15 markers / 71 reads and 13 markers / 432 reads carry data
    assert golden_problem(RESULT_PILEUP).used_counts() == (15, 71)
    assert golden_problem(LONGREAD_PILEUP).used_counts() == (13, 432)
    assert golden_problem(RESULT_PILEUP).n_marker_total == 9787


# reference ctest suite (CMakeLists.txt:86-147): mode flags -> golden .Ancestry
GOLDEN_MODES = [
    ("result.Ancestry", RESULT_PILEUP, {}),                                             # myTest1-4
    ("test.LongRead.pileup.Ancestry", LONGREAD_PILEUP, {}),                             # myTest5-6
    ("test.WithinAncestry.Ancestry", RESULT_PILEUP, {"within_ancestry": True}),
    ("test.WithinAncestry.FixPC.Ancestry", RESULT_PILEUP, {"within_ancestry": True, "fix_pc": [0.034756, 0.0193]}),
    ("test.FixAlpha.Ancestry", RESULT_PILEUP, {"fix_alpha": 0.1}),
    ("test.HeterFixPC.Ancestry", RESULT_PILEUP, {"fix_pc": [0.034756, 0.0193]}),
]


@pytest.mark.parametrize("golden,pileup,kw", GOLDEN_MODES)
def test_golden_ancestry(golden, pileup, kw):
    r = golden_problem(pileup).optimize(**kw)
    text = vo.format_ancestry(r["pc_contam"], r["pc_intended"])
    assert text == open(os.path.join(EXPECTED, golden)).read()


def test_golden_selfsm_numbers():
    # resource/test/expected/*.selfSM: FREEMIX, FREELK1, FREELK0 (6 significant digits)
    r = golden_problem(RESULT_PILEUP).optimize()
    assert "%g" % min(r["alpha"], 1 - r["alpha"]) == "0.21808"
    assert "%g" % -r["llk1"] == "-19.0924" and "%g" % -r["llk0"] == "-40.8945"
    r = golden_problem(LONGREAD_PILEUP).optimize()
    assert "%g" % min(r["alpha"], 1 - r["alpha"]) == "0.419446"
    assert "%g" % -r["llk1"] == "-388.883" and "%g" % -r["llk0"] == "-424.863"


def test_full_run_values():
    # SURVEY.md section 8(c) full-run known answers
    p = golden_problem(RESULT_PILEUP)
    assert abs(p.optimize()["alpha"] - 0.2180797296) < 1e-9
    assert abs(p.optimize(within_ancestry=True)["alpha"] - 0.2306066397) < 1e-9
    assert abs(p.optimize(within_ancestry=True, fix_pc=[0.034756, 0.0193])["alpha"] - 0.2142759671) < 1e-9
    assert abs(p.optimize(fix_pc=[0.034756, 0.0193])["alpha"] - 0.08823492334) < 1e-9


def test_parse_rules():
    # SimplePileupViewer.cpp:711-746: indels skipped without a qual, '^' skips a char, '*'/'#' consume a qual
    seq, qual = vo.parse_pileup_seq_bases_only(b"^]..+2AC,-1g*a$#T", b"ABCDEFG")
    assert seq == b"..,aT" and qual == b"ABCEG"
    # long-read fixture: qualities above Phred 60 survive parsing ('{' = 90)
    p = golden_problem(LONGREAD_PILEUP)
    assert p.quals.max() == ord("{")


@pytest.mark.skipif(not vo.ref_available(), reason="oracle/_ref/vb2_ref not built (no /root/reference here)")
def test_oracle_matches_reference_binary(tmp_path):
    pts = tmp_path / "pts.txt"
    pts.write_text("".join("%r %r %r %r %r\n" % (a[0], a[1], b[0], b[1], al) for a, b, al in KAT_POINTS))
    for pileup in (RESULT_PILEUP, LONGREAD_PILEUP):
        recs = vo.run_ref(["--DisableSanityCheck", "--PileupFile", pileup, "--SVDPrefix", HAPMAP, "--NumPC", "2",
                           "--Output", str(tmp_path / "o"), "--EvalPoints", str(pts)])
        p = golden_problem(pileup)
        ref_llk = [r["llk"] for r in recs if r["phase"] == "eval"]
        assert ref_llk == [p.compute_mix_llks(*pt) for pt in KAT_POINTS]          # bit-identical
        opt = [r for r in recs if r["phase"] == "optimize"][0]
        mine = p.optimize()
        assert mine["alpha"] == opt["alpha"] and mine["evals"] == opt["evals"]
        assert mine["pc_contam"] == opt["pc_contam"] and mine["pc_intended"] == opt["pc_intended"]
        assert mine["llk1"] == opt["llk1"] and mine["llk0"] == opt["llk0"]


@pytest.mark.skipif(not vo.ref_available(), reason="oracle/_ref/vb2_ref not built (no /root/reference here)")
def test_known_af_stream_semantics_match_reference_binary(tmp_path):
    """--KnownAF rows with a multi-allelic ALT ("A,G") make the reference's `ss >> alt >> AF` fail and use
    AF = 0 (ContaminationEstimator.cpp:476-484); the oracle's reader must do the same."""
    rows = []
    with open(HAPMAP + ".bed") as f:
        for i, line in enumerate(f):
            c, _, p, r, a = line.split()[:5]
            alt = a + ",T" if i % 7 == 0 else a
            rows.append("%s\t%d\t%s\t%s\t%s\t%r\n" % (c, int(p) - 1, p, r, alt, 0.05 + 0.009 * (i % 100)))
    af = tmp_path / "af.txt"
    af.write_text("".join(rows))
    pts = tmp_path / "pts.txt"
    pts.write_text("0 0 0 0 0.03\n0 0 0 0 0.3\n")
    recs = vo.run_ref(["--DisableSanityCheck", "--PileupFile", RESULT_PILEUP, "--SVDPrefix", HAPMAP, "--NumPC", "2",
                       "--KnownAF", str(af), "--Output", str(tmp_path / "o"), "--EvalPoints", str(pts)])
    p = vo.problem_from_files(HAPMAP, RESULT_PILEUP, 2, disable_sanity=True, known_af_path=str(af))
    assert (p.known_af == 0).sum() > 0
    assert [r["llk"] for r in recs if r["phase"] == "eval"] == [p.compute_mix_llks([0, 0], [0, 0], a) for a in (0.03, 0.3)]
    opt = [r for r in recs if r["phase"] == "optimize"][0]
    assert p.optimize()["alpha"] == opt["alpha"]
