"""BASELINE.json configs[2], [3] and [4] at FULL size against the CPU oracle (configs[1] is bench.py's headline and
`test_full_size_invariants`):
  configs[2]  1000g.phase3.100k panel, 30x, NumPC=4           -- FixedLayout<4>, simplex dimension 9
  configs[4]  hgdp.100k panel, 200x WGS depth, NumPC=4        -- ~20 M reads, the chunked (multi-stage) kernels
  configs[3]  64 samples x 100k markers x 30x                 -- one many-samples launch vs 64 single evaluations,
              and the lock-step cohort CLI on 64 samples vs 64 separate runs
Tolerances: LLK <= 1e-8 relative (fp32 panel in HBM), alpha / PCs within 1e-4 (north_star)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import to_oracle   # (puts the repo root on sys.path)
import bench
import verifybamid_b200 as vb
from verifybamid_b200 import host, panels, synth

pytestmark = pytest.mark.gpu
REL = 1e-8
TOL = 1e-4


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def _cli(args, cwd, env=None):
    cp = subprocess.run([host.CLI_PATH, *args], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                        timeout=1800)
    assert cp.returncode == 0, cp.stderr[-2000:]
    return cp


def _ancestry(path):
    rows = [l.split("\t") for l in open(path).read().splitlines()[1:]]
    return np.array([[float(r[1]), float(r[2])] for r in rows])


def _full_size_case(tmp_path, name, expect_chunked):
    s = bench.make_workload(name)
    p = s.problem
    k = p.n_pc
    ora = to_oracle(p, num_thread=min(32, os.cpu_count() or 1))
    pts = [([0.01] * k, [0.01] * k, 0.03), (list(s.pc_contam), list(s.pc_intended), 0.02),
           ([0.03, -0.02, 0.01, 0.005][:k], list(s.pc_intended), 0.3)]
    with vb.LLKEngine(p) as eng:
        info = eng.info()
        assert (info["markers_used"], info["reads_used"]) == ora.used_counts()
        single = [eng.compute_mix_llks(*pt) for pt in pts]
        for got, pt in zip(single, pts):
            assert rel(got, ora.compute_mix_llks(*pt)) <= REL
        batch = eng.eval_batch(np.array([pt[0] for pt in pts] * 3), np.array([pt[1] for pt in pts] * 3),
                               np.array([pt[2] for pt in pts] * 3))
        assert batch.tolist() == single * 3                       # the many-evaluations kernel: same bits
        # the deep sample does not fit on chip: no evaluation session (and hence no search on the device) for it
        if expect_chunked:
            with pytest.raises(vb.VB2Error):
                eng.session_begin()
        else:
            eng.session_begin()
            assert [eng.compute_mix_llks(*pt) for pt in pts] == single
            eng.session_end()
    # the full optimisation through the product CLI against the oracle's OptimizeLLK on the same problem
    prefix = panels.write_text_panel(s.panel, str(tmp_path / "panel"))
    pile = s.write_pileup(str(tmp_path / "sample.pileup"))
    out = str(tmp_path / "o")
    cp = _cli(["--SVDPrefix", prefix, "--PileupFile", pile, "--Reference", "x", "--NumPC", str(k), "--Output", out], str(tmp_path))
    assert ("Simplex search on the device" in cp.stderr) == (not expect_chunked)
    prob, _ = host.load_problem(prefix, pile, k, disable_sanity=False)
    want = to_oracle(prob, num_thread=min(32, os.cpu_count() or 1)).optimize()
    got = _ancestry(out + ".Ancestry")
    assert np.abs(got[:, 0] - np.array(want["pc_contam"])).max() <= TOL
    assert np.abs(got[:, 1] - np.array(want["pc_intended"])).max() <= TOL
    sm = open(out + ".selfSM").read().splitlines()[1].split("\t")
    assert abs(float(sm[6]) - min(want["alpha"], 1 - want["alpha"])) <= TOL
    assert abs(float(sm[7]) + want["llk1"]) <= 1e-6 * abs(want["llk1"])
    assert abs(float(sm[6]) - 0.02) < 5e-3                        # and it recovers the simulated contamination


def test_config2_100k_numpc4_full_size(tmp_path):
    _full_size_case(tmp_path, "k4", expect_chunked=False)


def test_config4_hgdp_200x_numpc4_full_size(tmp_path):
    _full_size_case(tmp_path, "hgdp200", expect_chunked=True)


def test_config3_64_samples_one_launch_full_size():
    panel = panels.load_bundled("1000g.phase3.100k.b37")
    n = 64
    rng = np.random.default_rng(7)
    samples = [synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=1 + i, sanity_check=True) for i in range(n)]
    engines = [vb.LLKEngine(s.problem, batched=True) for s in samples]
    try:
        pc1 = rng.normal(0.0, 0.02, (n, 2)); pc2 = rng.normal(0.0, 0.02, (n, 2)); al = rng.uniform(0.001, 0.4, n)
        got = vb.eval_many(engines, pc1, pc2, al)
        want = [e.compute_mix_llks(pc1[j], pc2[j], al[j]) for j, e in enumerate(engines)]
        assert got.tolist() == want                               # one launch for the cohort == 64 single evaluations
        for j in (0, 31, 63):
            ora = to_oracle(samples[j].problem, num_thread=min(32, os.cpu_count() or 1))
            assert rel(got[j], ora.compute_mix_llks(list(pc1[j]), list(pc2[j]), float(al[j]))) <= REL
    finally:
        for e in engines:
            e.close()


def test_config3_cohort_cli_64_samples(tmp_path):
    """--PileupList with 64 samples (10k panel, so that 64 pileup texts stay small): the lock-step cohort writes exactly
    the files 64 separate runs write."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    prefix = None
    lines, singles = [], []
    for i in range(64):
        s = synth.make_sample(panel, n_pc=2, depth=[30.0, 12.0, 45.0, 20.0][i % 4], alpha=[0.02, 0.1, 0.005, 0.3][(i // 4) % 4],
                              seed=500 + i, n_markers=4000)
        if prefix is None:
            prefix = panels.write_text_panel(s.panel, str(tmp_path / "panel"))
        pile = s.write_pileup(str(tmp_path / ("s%d.pileup" % i)))
        lines.append("%s\t%s\n" % (pile, tmp_path / ("cohort%d" % i)))
        singles.append((pile, str(tmp_path / ("single%d" % i))))
    lst = tmp_path / "cohort.list"
    lst.write_text("".join(lines))
    cp = _cli(["--PileupList", str(lst), "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2"], str(tmp_path))
    table = [l.split("\t") for l in cp.stdout.splitlines() if l and not l.startswith("#")]
    assert len(table) == 64 and all(row[-1] == "OK" for row in table)
    for i, (pile, out) in enumerate(singles):
        if i % 8:                                                  # every eighth sample also on its own
            continue
        _cli(["--PileupFile", pile, "--SVDPrefix", prefix, "--Reference", "x", "--NumPC", "2", "--Output", out], str(tmp_path))
        for ext in (".Ancestry", ".selfSM"):
            assert open(out + ext).read() == open(str(tmp_path / ("cohort%d" % i)) + ext).read(), (i, ext)
