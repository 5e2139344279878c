"""Host flatten (llk_pack.cpp) checked on the CPU: the packed image, run through a numpy
restatement of the kernel's arithmetic, must reproduce the oracle."""
import numpy as np
import pytest

import verifybamid_b200 as vb
from verifybamid_b200 import panels, synth
from helpers import (KAT_POINTS, KAT_RESULT, KAT_LONGREAD, LONGREAD_PILEUP, RESULT_PILEUP, emulate_packed_llk, iter_blobs,
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
    assert pk["n_slices"] == 1 and pk["grid_x"] == 1 and pk["n_bins"] == 4 and pk["n_rounds"] == 1
    # padding lanes are marked and carry no reads
    assert (pk["marker_index"][13:] == 0xFFFFFFFF).all()
    (_, _, blob), = list(iter_blobs(pk))
    wr, wa, nv, _ = blob[:16].view(np.uint32)
    n_valid = int(nv) & 0xFF
    assert n_valid == 13 and blob.size == pk["off_words"] + (wr + wa) * 128 == pk["rounds"][0]["stride"]
    byts = blob[pk["off_words"]:]
    assert ((byts <= 93) | (byts == 0xFF)).all()
    assert int((byts != 0xFF).sum()) == pk["reads_streamed"]


def test_rounds_are_balanced_and_aligned():
    """Dealing by cost (costliest slice of a round to the least loaded bin): every SM sub-partition bin gets (nearly)
    the same number of word rows; every blob is 16-byte aligned (TMA bulk copy) and its size a multiple of 16."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    p = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=3).problem
    pk = vb.pack_host(p, max_ctas=8)                       # 32 bins, 313 slices -> 10 rounds
    assert pk["n_bins"] == 32 and pk["n_rounds"] == -(-pk["n_slices"] // 32) and pk["conc_rounds"] == 4
    work = np.zeros(pk["n_bins"])
    seen = set()
    for j, b, blob in iter_blobs(pk):
        wr, wa = blob[:8].view(np.uint32)
        work[b] += wr + wa
        assert (j // pk["n_bins"], b) not in seen           # one blob per (round, bin)
        seen.add((j // pk["n_bins"], b))
    assert work.max() <= 1.06 * work.mean()
    for R in pk["rounds"]:
        assert R["base"] % 16 == 0 and R["stride"] % 16 == 0
        assert R["stride"] == pk["off_words"] + 128 * R["rows"]
        assert R["first_bin"] + R["count"] == pk["n_bins"] or R["first_bin"] == 0
    strides = [R["stride"] for R in pk["rounds"]]
    assert strides == sorted(strides, reverse=True)         # heaviest slices first


@pytest.mark.parametrize("q_lo,q_hi", [(20, 40), (0, 12)])
def test_kernel_arithmetic_form_matches_oracle(q_lo, q_hi):
    """The form the kernels run (pairs of reads through symmetric functions, log of the product of a bin's marginals)
    against the oracle, including Phred 0..12 where the pair products cancel the most -- on the CPU, in numpy."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=25.0, alpha=0.05, seed=21, n_markers=3000, q_lo=q_lo, q_hi=q_hi)
    ora = to_oracle(s.problem)
    pk = vb.pack_host(s.problem, max_ctas=8)
    for pc1, pc2, a in [([0.01, 0.01], [0.01, 0.01], 0.03), ([0.02, -0.01], [-0.01, 0.027], 0.05),
                        ([0.0, 0.0], [0.0, 0.0], 0.5), ([0.01, 0.01], [0.01, 0.01], 1e-6), ([0.05, -0.02], [0.01, 0.01], 0.999)]:
        want = ora.compute_mix_llks(pc1, pc2, a)
        assert abs(emulate_packed_llk(pk, pc1, pc2, a, kernel_form=True) - want) <= 1e-11 * abs(want)
        assert abs(emulate_packed_llk(pk, pc1, pc2, a) - want) <= 1e-11 * abs(want)


def test_deal_is_level_by_cost_and_batched_layout_is_deeper():
    """The deal is by what the kernel spends (4 per full row, 2 per uniform tail, 7 per checked row, 15 per slice, in
    quarter rows); VB2_FLAG_BATCHED lays a small shard out over fewer bins with >= 5 slices each -- same likelihood."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=3)
    pk = vb.pack_host(s.problem, max_ctas=8)
    cost = np.zeros(pk["n_bins"])
    for _, b, blob in iter_blobs(pk):
        wr, wa, z, w = (int(x) for x in blob[:16].view(np.uint32))
        fr, fa, tr, ta = w & 0xFFFF, w >> 16, (z >> 8) & 0xF, (z >> 12) & 0xF
        rr, ra = wr - fr, wa - fa
        cost[b] += 4 * (fr + fa) + (2 if (tr and rr == 1) else 7 * rr) + (2 if (ta and ra == 1) else 7 * ra) + 15
    assert cost.max() <= 1.03 * cost.mean()
    wide = vb.pack_host(s.problem, shard_rank=1, shard_count=8)          # 1/8 of 313 slices over 148 SMs' bins
    deep = vb.pack_host(s.problem, shard_rank=1, shard_count=8, batched=True)
    assert wide["n_rounds"] == 1 and deep["n_rounds"] >= 5 and deep["n_bins"] < wide["n_bins"]
    pt = ([0.02, -0.01], [-0.01, 0.027], 0.05)
    a, b = emulate_packed_llk(wide, *pt), emulate_packed_llk(deep, *pt)
    assert abs(a - b) <= 1e-12 * abs(a)


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
    # the same sample dealt to a small launch (many rounds per bin) gives the same likelihood
    pk8 = vb.pack_host(p, max_ctas=3)
    assert pk8["n_rounds"] > 2
    assert abs(emulate_packed_llk(pk8, [0.02, -0.01], [-0.01, 0.027], 0.05)
               - emulate_packed_llk(pk, [0.02, -0.01], [-0.01, 0.027], 0.05)) <= 1e-9


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
    assert pk["n_used"] == 0 and pk["n_slices"] == 0 and pk["blob"].size == 0 and pk["n_rounds"] == 0
    assert emulate_packed_llk(pk, [0, 0], [0, 0], 0.5) == 0.0   # reference: empty sum (h:231)
