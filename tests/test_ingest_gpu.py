"""Device-side pileup ingest (llk_ingest.cu: vb2_panel_create / vb2_ingest_parse / vb2_ingest_flatten) against the host
path it replaces: the C++ host reader (SimplePileupViewer.cpp:711-833 rules) + BuildResolvedMarkers
(ContaminationEstimator.cpp:67-86) + the host flatten.  The bar: the image in device memory is BYTE-IDENTICAL to the one
vb2_llk_pack_host builds from the host reader's arrays, and text with the reference's parsing quirks is refused
(VB2_ERR_UNSUPPORTED) instead of being parsed differently."""
import os

import numpy as np
import pytest

from helpers import GOLD, HAPMAP, LONGREAD_PILEUP, RESULT_PILEUP, to_oracle
import bench
import verifybamid_b200 as vb
from verifybamid_b200 import host, panels, synth

pytestmark = pytest.mark.gpu


def _device_panel(panel, n_pc):
    return vb.DevicePanel(panel.ud[:, :n_pc], panel.mu, panel.chrom, panel.pos, panel.alt_char())


def _compare(tmp_path, s, n_pc, sanity_disabled):
    prefix = panels.write_text_panel(s.panel, str(tmp_path / "panel"))
    pile = s.write_pileup(str(tmp_path / "sample.pileup"))
    prob, summ = host.load_problem(prefix, pile, n_pc, disable_sanity=sanity_disabled)   # the host reader + sanity check
    dp = _device_panel(s.panel, n_pc)
    text = open(pile, "rb").read()
    eng = dp.ingest(text, sanity_disabled=sanity_disabled)
    try:
        info = eng.info()
        want = vb.pack_host(prob, max_ctas=info["sm_count"], panel_dtype=vb.VB2_PANEL_FP32)
        # what the host's sanity check is fed: the same depths, the same mean / sd
        assert eng.ingest_summary["avg_depth"] == prob.avg_depth
        if not sanity_disabled:
            assert eng.ingest_summary["sd_depth"] == prob.sd_depth
        assert np.array_equal(eng.ingest_summary["row_depth"].clip(min=0), prob.depths())
        # the image: sizes, constants, and every byte
        assert info["markers_used"] == want["n_used"] and info["n_slices"] == want["n_slices"] and info["grid_x"] == want["grid_x"]
        assert (info["reads_used"], info["reads_streamed"], info["reads_folded"]) == (want["reads_used"], want["reads_streamed"], want["reads_folded"])
        assert info["log_other_const"] == want["log_other_const"]
        assert info["device_bytes"] == want["blob_bytes"]
        got = eng.debug_image(want["blob_bytes"])
        assert np.array_equal(got, want["blob"])
        # ... and therefore the same likelihood, bit for bit, as the context built from the host reader's arrays
        with vb.LLKEngine(prob) as ref:
            for pc1, pc2, a in (([0.01] * n_pc, [0.01] * n_pc, 0.03), (list(s.pc_contam), list(s.pc_intended), 0.02)):
                assert eng.compute_mix_llks(pc1, pc2, a) == ref.compute_mix_llks(pc1, pc2, a)
    finally:
        eng.close()
        dp.close()


def test_image_is_byte_identical_10k(tmp_path):
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=4)
    _compare(tmp_path, s, 2, sanity_disabled=False)


def test_image_is_byte_identical_numpc4_sanity_disabled(tmp_path):
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=4, depth=12.0, alpha=0.1, seed=9, n_markers=6000, q_lo=0, q_hi=45)
    _compare(tmp_path, s, 4, sanity_disabled=True)


def test_image_is_byte_identical_headline_workload(tmp_path):
    _compare(tmp_path, bench.make_workload("100k30x"), 2, sanity_disabled=False)


def _golden_panel():
    ud = np.loadtxt(HAPMAP + ".UD", ndmin=2)
    mu = np.array([float(l.split()[1]) for l in open(HAPMAP + ".mu")])
    bed = [l.split() for l in open(HAPMAP + ".bed")]
    return ud, mu, [b[0] for b in bed], np.array([int(b[2]) for b in bed]), np.frombuffer("".join(b[4][0] for b in bed).encode(), np.uint8)


@pytest.mark.parametrize("pileup", [RESULT_PILEUP, LONGREAD_PILEUP])
def test_reference_fixtures(pileup):
    """The reference's own pileups (indels, read starts/ends, deletions, long reads) through the device reader."""
    ud, mu, chrom, pos, alt = _golden_panel()
    prob, _ = host.load_problem(HAPMAP, pileup, 2, disable_sanity=True)
    dp = vb.DevicePanel(ud[:, :2], mu, chrom, pos, alt)
    eng = dp.ingest(open(pileup, "rb").read(), sanity_disabled=True, panel_dtype=vb.VB2_PANEL_FP64)
    try:
        want = vb.pack_host(prob, max_ctas=eng.info()["sm_count"], panel_dtype=vb.VB2_PANEL_FP64)
        assert np.array_equal(eng.ingest_summary["row_depth"].clip(min=0), prob.depths())
        assert np.array_equal(eng.debug_image(want["blob_bytes"]), want["blob"])
        ora = to_oracle(prob)
        for pc1, pc2, a in (([0.0, 0.0], [0.0, 0.0], 0.5), ([0.01, 0.01], [0.01, 0.01], 0.03)):
            got, ref = eng.compute_mix_llks(pc1, pc2, a), ora.compute_mix_llks(pc1, pc2, a)
            assert abs(got - ref) <= 1e-11 * abs(ref)
    finally:
        eng.close()
        dp.close()


def test_quirky_text_is_left_to_the_host_reader():
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=8.0, alpha=0.02, seed=2, n_markers=300)
    dp = _device_panel(s.panel, 2)
    c, p0, p1 = s.panel.chrom[0], int(s.panel.pos[0]), int(s.panel.pos[1])
    good = ("%s\t%d\tA\t3\t.,.\tIII\n%s\t%d\tC\t2\t.$^F,\tII\n" % (c, p0, c, p1)).encode()
    try:
        dp.ingest(good, sanity_disabled=True).close()                       # well-formed: accepted
        dp.ingest(good[:-1], sanity_disabled=True).close()                  # ... also without the final newline
        for bad in (good + b"\n",                                           # an empty line
                    good + ("%s\t%d\tA\t3\n" % (c, p0 + 7)).encode(),       # a short line
                    good + ("%s\t12x\tA\t3\t...\tIII\n" % c).encode(),      # a position that is not a number
                    good + ("%s\t%d\tA\t1\t.\tI\n" % (c, p0)).encode(),     # a duplicated position
                    good + ("%s\t%d\t.\t1\t.\tI\n" % (c, p0 + 9)).encode(), # '.' reference next to '.' bases
                    good + ("%s\t%d\tA\t2\t.+x\tII\n" % (c, p0 + 9)).encode(),   # indel without a length
                    good + ("%s\t%d\tA\t3\t...\tII\n" % (c, p0 + 9)).encode()):  # fewer qualities than bases
            with pytest.raises(vb.VB2Error) as e:
                dp.ingest(bad, sanity_disabled=True)
            assert e.value.code == vb.VB2_ERR_UNSUPPORTED, bad
    finally:
        dp.close()
