"""End to end through the C++ host (verifybamid_b200/VerifyBamID): reference flags in, .Ancestry/.selfSM out.
Mirrors the reference's ctest suite (CMakeLists.txt:86-147): every model flag combination against the golden
files of resource/test/expected, plus synthetic panels against the oracle's full optimisation.

Tolerance: BASELINE.json north_star -- final alpha and PCs within 1e-4 of the reference CPU path."""
import os
import subprocess

import numpy as np
import pytest

import verifybamid_b200 as vb
from verifybamid_b200 import host, panels, synth
from helpers import GOLD, HAPMAP, LONGREAD_PILEUP, RESULT_PILEUP, golden_problem, to_oracle, vo

pytestmark = pytest.mark.gpu
TOL = 1e-4
EXPECTED = os.path.join(GOLD, "expected")


def run_cli(args, cwd):
    cp = subprocess.run([host.CLI_PATH, *args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                        timeout=600)
    assert cp.returncode == 0, cp.stderr[-2000:]
    return cp


def read_ancestry(path):
    rows = [l.split("\t") for l in open(path).read().splitlines()]
    assert rows[0] == ["PC", "ContaminatingSample", "IntendedSample"]
    return np.array([[float(r[1]), float(r[2])] for r in rows[1:]])


def read_selfsm(path):
    head, row = [l.split("\t") for l in open(path).read().splitlines()]
    return dict(zip(head, row))


MODES = [
    ("result.Ancestry", RESULT_PILEUP, []),
    ("test.LongRead.pileup.Ancestry", LONGREAD_PILEUP, []),
    ("test.WithinAncestry.Ancestry", RESULT_PILEUP, ["--WithinAncestry"]),
    ("test.WithinAncestry.FixPC.Ancestry", RESULT_PILEUP, ["--WithinAncestry", "--FixPC", "0.034756:0.0193"]),
    ("test.FixAlpha.Ancestry", RESULT_PILEUP, ["--FixAlpha", "0.1"]),
    ("test.HeterFixPC.Ancestry", RESULT_PILEUP, ["--FixPC", "0.034756:0.0193"]),
]


@pytest.mark.parametrize("golden,pileup,flags", MODES)
@pytest.mark.parametrize("panel_flag", [[], ["--PanelFP64"]])
def test_reference_ctest_goldens(tmp_path, golden, pileup, flags, panel_flag):
    out = str(tmp_path / "r")
    run_cli(["--DisableSanityCheck", "--PileupFile", pileup, "--SVDPrefix", HAPMAP, "--Reference", "chr20.fa.gz",
             "--NumPC", "2", "--Output", out, *flags, *panel_flag], str(tmp_path))
    got = read_ancestry(out + ".Ancestry")
    want = read_ancestry(os.path.join(EXPECTED, golden))
    assert np.abs(got - want).max() <= TOL, (got, want)
    if panel_flag:                       # fp64 panel: the 6-significant-digit text is reproduced exactly
        assert open(out + ".Ancestry").read() == open(os.path.join(EXPECTED, golden)).read()


def test_selfsm_matches_reference_goldens(tmp_path):
    for pileup, golden in ((RESULT_PILEUP, "result.selfSM"), (LONGREAD_PILEUP, "test.LongRead.pileup.selfSM")):
        out = str(tmp_path / "s")
        run_cli(["--DisableSanityCheck", "--PileupFile", pileup, "--SVDPrefix", HAPMAP, "--Reference", "x", "--NumPC", "2",
                 "--Output", out], str(tmp_path))
        got, want = read_selfsm(out + ".selfSM"), read_selfsm(os.path.join(EXPECTED, golden))
        assert list(got) == list(want)
        for key in ("#SEQ_ID", "RG", "CHIP_ID", "#SNPS", "AVG_DP"):
            assert got[key] == want[key]
        assert got["#READS"] == "NA"                                   # pileup input (main.cpp:398-400)
        for key in ("FREEMIX", "FREELK1", "FREELK0"):
            assert abs(float(got[key]) - float(want[key])) <= 1e-4 * max(1.0, abs(float(want[key])))


@pytest.fixture(scope="module")
def synthetic10k(tmp_path_factory):
    td = tmp_path_factory.mktemp("syn")
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=4)
    prefix = panels.write_text_panel(s.panel, str(td / "panel"))
    pile = s.write_pileup(str(td / "sample.pileup"))
    return s, prefix, pile, td


@pytest.mark.parametrize("flags,kw", [([], {}), (["--WithinAncestry"], {"within_ancestry": True}),
                                       (["--FixAlpha", "0.05"], {"fix_alpha": 0.05})])
def test_synthetic_10k_converges_to_the_oracle(synthetic10k, flags, kw):
    s, prefix, pile, td = synthetic10k
    out = str(td / ("o" + "".join(flags).replace("-", "")))
    cp = run_cli(["--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2", "--Output", out, *flags],
                 str(td))
    # the oracle runs the same model on the problem the C++ host parsed (sanity filter on: CLI default)
    prob, summ = host.load_problem(prefix, pile, 2, disable_sanity=False)
    want = to_oracle(prob).optimize(**kw)
    got = read_ancestry(out + ".Ancestry")
    assert np.abs(got[:, 0] - np.array(want["pc_contam"])).max() <= TOL
    assert np.abs(got[:, 1] - np.array(want["pc_intended"])).max() <= TOL
    sm = read_selfsm(out + ".selfSM")
    assert abs(float(sm["FREEMIX"]) - min(want["alpha"], 1 - want["alpha"])) <= TOL
    assert abs(float(sm["FREELK1"]) + want["llk1"]) <= 1e-6 * abs(want["llk1"])
    assert sm["#SNPS"] == str(prob.n_marker)
    if not flags:
        assert abs(float(sm["FREEMIX"]) - 0.02) < 5e-3               # and it recovers the simulated truth
        assert "Estimation from OptimizeHeter:" in cp.stdout and "FREEMIX(Alpha):" in cp.stdout


def test_four_pcs_and_output_pileup_round_trip(tmp_path):
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=4, depth=25.0, alpha=0.05, seed=8, n_markers=6000)
    prefix = panels.write_text_panel(s.panel, str(tmp_path / "p"))
    pile = s.write_pileup(str(tmp_path / "s.pileup"))
    out = str(tmp_path / "o")
    run_cli(["--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "4", "--Output", out,
             "--OutputPileup"], str(tmp_path))
    prob, _ = host.load_problem(prefix, pile, 4, disable_sanity=False)
    want = to_oracle(prob).optimize()
    got = read_ancestry(out + ".Ancestry")
    assert got.shape == (4, 2)
    assert np.abs(got[:, 0] - np.array(want["pc_contam"])).max() <= TOL
    assert np.abs(got[:, 1] - np.array(want["pc_intended"])).max() <= TOL
    assert abs(float(read_selfsm(out + ".selfSM")["FREEMIX"]) - min(want["alpha"], 1 - want["alpha"])) <= TOL
    # --OutputPileup is a cache of the input stage (main.cpp:336-369): feeding it back gives the same estimate
    out2 = str(tmp_path / "o2")
    run_cli(["--PileupFile", out + ".Pileup", "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "4", "--Output", out2],
            str(tmp_path))
    assert open(out2 + ".Ancestry").read() == open(out + ".Ancestry").read()


def test_cli_errors_like_the_reference(tmp_path):
    cp = subprocess.run([host.CLI_PATH, "--SVDPrefix", HAPMAP, "--PileupFile", RESULT_PILEUP], cwd=str(tmp_path),
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode != 0 and "--Reference is required" in cp.stderr
    cp = subprocess.run([host.CLI_PATH, "--SVDPrefix", HAPMAP, "--Reference", "x", "--BamFile", "a.bam"], cwd=str(tmp_path),
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode != 0 and "htslib" in cp.stderr
    cp = subprocess.run([host.CLI_PATH, "--SVDPrefix", HAPMAP, "--Reference", "x", "--PileupFile", RESULT_PILEUP,
                         "--NumPC", "9"], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert cp.returncode != 0 and "--NumPC" in cp.stderr               # only 2 PCs in the hapmap fixture


@pytest.mark.skipif(vb.device_count() < 2, reason="needs two GPUs")
def test_marker_shards_over_two_gpus(synthetic10k):
    s, prefix, pile, td = synthetic10k
    outs = []
    for n in (1, 2):
        out = str(td / ("g%d" % n))
        run_cli(["--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2", "--Output", out,
                 "--NumGPU", str(n)], str(td))
        outs.append(read_ancestry(out + ".Ancestry"))
    assert np.abs(outs[0] - outs[1]).max() <= TOL


def test_cohort_mode_equals_one_sample_at_a_time(tmp_path):
    """--PileupList (BASELINE configs[3] shape, scaled down): five samples optimised in lock-step, one launch per
    simplex step for the whole cohort, must write exactly the files five separate runs write."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    prefix = None
    lines, singles = [], []
    for i, (depth, alpha, nm) in enumerate([(30.0, 0.02, 5000), (12.0, 0.10, 5000), (45.0, 0.005, 5000),
                                            (20.0, 0.3, 5000), (30.0, 0.02, 5000)]):
        s = synth.make_sample(panel, n_pc=2, depth=depth, alpha=alpha, seed=50 + i, n_markers=nm)
        if prefix is None:
            prefix = panels.write_text_panel(s.panel, str(tmp_path / "panel"))
        pile = s.write_pileup(str(tmp_path / ("s%d.pileup" % i)))
        lines.append("%s\t%s\n" % (pile, tmp_path / ("cohort%d" % i)))
        singles.append((pile, str(tmp_path / ("single%d" % i))))
    lst = tmp_path / "cohort.list"
    lst.write_text("".join(lines))
    cp = run_cli(["--PileupList", str(lst), "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2"], str(tmp_path))
    table = [l.split("\t") for l in cp.stdout.splitlines() if l and not l.startswith("#")]
    assert len(table) == 5 and all(row[-1] == "OK" for row in table)
    assert "launches" in cp.stderr
    for i, (pile, out) in enumerate(singles):
        run_cli(["--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2", "--Output", out],
                str(tmp_path))
        for ext in (".Ancestry", ".selfSM"):
            assert open(out + ext).read() == open(str(tmp_path / ("cohort%d" % i)) + ext).read(), (i, ext)


def _evals(stderr: str):
    import re
    m = re.search(r"Likelihood evaluations: (\d+)", stderr)
    d = re.search(r"Simplex search on the device: (\d+)", stderr)
    return int(m.group(1)), int(d.group(1)) if d else 0


@pytest.mark.parametrize("golden,pileup,flags", MODES)
def test_device_simplex_follows_the_host_simplex(tmp_path, golden, pileup, flags):
    """vb2_llk_minimize (AmoebaMinimizer::Minimize on the device, MathGenMin.cpp:326-423) against the same search driven
    evaluation by evaluation from the host (VB2_HOST_SIMPLEX=1): same number of evaluations, same files, all six models."""
    args = ["--DisableSanityCheck", "--PileupFile", pileup, "--SVDPrefix", HAPMAP, "--Reference", "x", "--NumPC", "2", *flags]
    dev = run_cli([*args, "--Output", str(tmp_path / "dev")], str(tmp_path))
    env = dict(os.environ, VB2_HOST_SIMPLEX="1")
    hostrun = subprocess.run([host.CLI_PATH, *args, "--Output", str(tmp_path / "host")], cwd=str(tmp_path), env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert hostrun.returncode == 0, hostrun.stderr[-2000:]
    n_dev, on_dev = _evals(dev.stderr)
    n_host, on_dev_host = _evals(hostrun.stderr)
    assert on_dev > 0 and on_dev_host == 0          # the search really ran where it should
    assert n_dev == n_host, (n_dev, n_host)
    for ext in (".Ancestry", ".selfSM"):
        assert open(str(tmp_path / "dev") + ext).read() == open(str(tmp_path / "host") + ext).read(), ext
    assert dev.stdout == hostrun.stdout


def test_device_simplex_on_synthetic_10k(synthetic10k):
    s, prefix, pile, td = synthetic10k
    outs = {}
    for name, env in (("dev", os.environ), ("host", dict(os.environ, VB2_HOST_SIMPLEX="1"))):
        out = str(td / ("simplex_" + name))
        cp = subprocess.run([host.CLI_PATH, "--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2",
                             "--Output", out], cwd=str(td), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                            timeout=600)
        assert cp.returncode == 0, cp.stderr[-2000:]
        outs[name] = (_evals(cp.stderr)[0], open(out + ".Ancestry").read(), open(out + ".selfSM").read())
    assert outs["dev"] == outs["host"]
