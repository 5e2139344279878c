"""Host flatten (llk_pack.cpp) checked on the CPU: the packed image, run through a numpy
restatement of the kernel's arithmetic, must reproduce the oracle."""
import numpy as np
import pytest

import verifybamid_b200 as vb
from verifybamid_b200 import panels, synth
from helpers import (KAT_POINTS, KAT_RESULT, KAT_LONGREAD, LONGREAD_PILEUP, RESULT_PILEUP, emulate_packed_llk,
                     golden_problem, to_oracle, to_product)


@pytest.mark.parametrize("pileup,kat", [(RESULT_PILEUP, KAT_RESULT), (LONGREAD_PILEUP, KAT_LONGREAD)])
def test_packed_image_reproduces_known_answers(pileup, kat):
    pk = vb.pack_host(to_product(golden_problem(pileup)))
    for (pc1, pc2, a), want in zip(KAT_POINTS, kat):
        got = emulate_packed_llk(pk, pc1, pc2, a)
        assert abs(got - want) <= 1e-11 * abs(want), (got, want)


def test_pack_counts_and_layout():
    p = to_product(golden_problem(LONGREAD_PILEUP))
    pk = vb.pack_host(p)
    assert pk["n_used"] == 13 and pk["reads_used"] == 432
    assert pk["reads_streamed"] + pk["reads_folded"] == pk["reads_used"]
    assert pk["n_slices"] == 1 and pk["m_pad"] == 32
    # padding lanes are marked and carry no reads
    assert (pk["marker_index"][13:] == 0xFFFFFFFF).all()
    wr, wa = pk["slice_desc"][0, 1] & 0xFFFF, pk["slice_desc"][0, 1] >> 16
    assert pk["words"].size == (wr + wa) * 32
    byts = pk["words"].view(np.uint8)
    assert ((byts <= 93) | (byts == 0xFF)).all()
    assert int((byts != 0xFF).sum()) == pk["reads_streamed"]


@pytest.fixture(scope="module")
def small_sample():
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    return synth.make_sample(panel, n_pc=2, depth=12.0, alpha=0.05, seed=11, n_markers=1500)


def test_packed_synthetic_matches_oracle(small_sample):
    p = small_sample.problem
    ora = to_oracle(p)
    pk = vb.pack_host(p)
    assert (pk["n_used"], pk["reads_used"]) == ora.used_counts() == p.used_counts()
    for pc1, pc2, a in ([[0.01, 0.01], [0.01, 0.01], 0.03], [[0.02, -0.01], [-0.01, 0.027], 0.05],
                        [[0.0, 0.0], [0.0, 0.0], 0.0]):
        want = ora.compute_mix_llks(pc1, pc2, a)
        got = emulate_packed_llk(pk, pc1, pc2, a)
        assert abs(got - want) <= 1e-11 * abs(want)


def test_sanity_filter_matches_reference_rule(small_sample):
    p = small_sample.problem          # generated with the +-3 sd depth window on
    assert not p.sanity_disabled and p.sd_depth > 0
    m_used, _ = p.used_counts()
    assert 0 < m_used < p.n_marker     # the window drops a few markers
    assert vb.pack_host(p)["n_used"] == m_used


@pytest.mark.parametrize("n_shards", [2, 4, 8])
def test_shards_partition_the_sample(small_sample, n_shards):
    p = small_sample.problem
    whole = vb.pack_host(p)
    parts = [vb.pack_host(p, r, n_shards) for r in range(n_shards)]
    assert sum(x["n_used"] for x in parts) == whole["n_used"]
    assert sum(x["reads_used"] for x in parts) == whole["reads_used"]
    rows = np.concatenate([x["marker_index"][x["marker_index"] != 0xFFFFFFFF] for x in parts])
    assert len(np.unique(rows)) == whole["n_used"]           # disjoint cover
    args = ([0.02, -0.01], [-0.01, 0.027], 0.05)
    total = sum(emulate_packed_llk(x, *args) for x in parts)
    assert abs(total - emulate_packed_llk(whole, *args)) <= 1e-10 * abs(total)


def test_empty_and_absent_markers():
    p = to_product(golden_problem(RESULT_PILEUP))
    # a sample with no usable marker flattens to nothing
    none = vb.PileupProblem(p.ud, p.means, np.full(p.n_marker, -1, np.int32), p.alt_base, np.zeros(1, np.int64),
                            np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    pk = vb.pack_host(none)
    assert pk["n_used"] == 0 and pk["n_slices"] == 0 and pk["words"].size == 0
    assert emulate_packed_llk(pk, [0, 0], [0, 0], 0.5) == 0.0   # reference: empty sum (h:231)
