"""Parity of the sm_100a path with the CPU oracle, through the C ABI (ctypes -> libvb2llk.so).

Tolerances (fp64 arithmetic on both sides, different operation order):
  REL_FP64  panel stored fp64 in HBM: only rounding-order differences          -> 1e-11 relative
  REL_FP32  panel stored fp32 (north-star layout): UD/mu rounded to fp32 once  -> 1e-8  relative
            (a fixed perturbation of the panel, smooth in the parameters; measured 5e-11 .. 1.3e-9)
(BASELINE north_star: final alpha and PCs within 1e-4; SURVEY section 7: LLK <= 1e-8 relative.)
"""
import os

import numpy as np
import pytest

import verifybamid_b200 as vb
from verifybamid_b200 import panels, synth
from helpers import (KAT_LONGREAD, KAT_POINTS, KAT_RESULT, LONGREAD_PILEUP, RESULT_PILEUP, golden_problem, to_oracle,
                     to_product)

pytestmark = pytest.mark.gpu

REL_FP64 = 1e-11
REL_FP32 = 1e-8
# the reference's own fixtures have 13-15 informative markers: the fp32 rounding of UD/mu is a random walk
# ~1e-7 * sqrt(markers) absolute, which on |LLK| ~ 20 is ~6e-9 relative (and 5e-11 at 4k markers)
REL_FP32_TINY = 5e-8

POINTS = [([0.01, 0.01], [0.01, 0.01], 0.03), ([0.02, -0.01], [-0.01, 0.027], 0.05), ([0.0, 0.0], [0.0, 0.0], 0.5),
          ([-0.0101, 0.0269], [-0.0101, 0.0269], 0.0), ([0.05, -0.02], [0.01, 0.01], 0.999)]


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


@pytest.mark.parametrize("pileup,kat", [(RESULT_PILEUP, KAT_RESULT), (LONGREAD_PILEUP, KAT_LONGREAD)])
@pytest.mark.parametrize("dtype", [vb.VB2_PANEL_FP64, vb.VB2_PANEL_FP32])
def test_reference_known_answers(pileup, kat, dtype):
    tol = REL_FP64 if dtype == vb.VB2_PANEL_FP64 else REL_FP32_TINY
    with vb.LLKEngine(to_product(golden_problem(pileup)), panel_dtype=dtype) as eng:
        for (pc1, pc2, a), want in zip(KAT_POINTS, kat):
            assert rel(eng.compute_mix_llks(pc1, pc2, a), want) <= tol


@pytest.fixture(scope="module")
def sample10k():
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    return synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=4)


@pytest.mark.parametrize("dtype,tol", [(vb.VB2_PANEL_FP64, REL_FP64), (vb.VB2_PANEL_FP32, REL_FP32)])
def test_synthetic_10k_matches_oracle(sample10k, dtype, tol):
    p = sample10k.problem
    ora = to_oracle(p)
    with vb.LLKEngine(p, panel_dtype=dtype) as eng:
        info = eng.info()
        assert (info["markers_used"], info["reads_used"]) == ora.used_counts()
        assert info["reads_streamed"] + info["reads_folded"] == info["reads_used"]
        for pc1, pc2, a in POINTS:
            assert rel(eng.compute_mix_llks(pc1, pc2, a), ora.compute_mix_llks(pc1, pc2, a)) <= tol


def test_four_pcs(sample10k):
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=4, depth=20.0, alpha=0.03, seed=9, n_markers=5000)
    ora = to_oracle(s.problem)
    with vb.LLKEngine(s.problem, panel_dtype=vb.VB2_PANEL_FP64) as eng:
        for pc in ([0.01] * 4, list(s.pc_intended), [0.03, -0.02, 0.01, 0.005]):
            assert rel(eng.compute_mix_llks(pc, list(s.pc_intended), 0.03),
                       ora.compute_mix_llks(pc, list(s.pc_intended), 0.03)) <= REL_FP64


def test_low_quality_reads_match_oracle():
    """Phred 0..12: e up to 1.0, where the pair products F(ea)F(eb) = C0 + C1(ea+eb) + C2 ea eb cancel the most."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=25.0, alpha=0.05, seed=21, n_markers=4000, q_lo=0, q_hi=12)
    ora = to_oracle(s.problem)
    with vb.LLKEngine(s.problem, panel_dtype=vb.VB2_PANEL_FP64) as eng:
        for pc1, pc2, a in POINTS + [([0.01, 0.01], [0.01, 0.01], 1e-6)]:
            assert rel(eng.compute_mix_llks(pc1, pc2, a), ora.compute_mix_llks(pc1, pc2, a)) <= REL_FP64
        got = eng.eval_batch(np.array([p[0] for p in POINTS]), np.array([p[1] for p in POINTS]),
                             np.array([p[2] for p in POINTS]))
        for g, (pc1, pc2, a) in zip(got, POINTS):
            assert rel(g, ora.compute_mix_llks(pc1, pc2, a)) <= REL_FP64


def test_bit_reproducible_and_batch_equals_single(sample10k):
    with vb.LLKEngine(sample10k.problem) as eng:
        single = [eng.compute_mix_llks(*pt) for pt in POINTS]
        again = [eng.compute_mix_llks(*pt) for pt in POINTS]
        assert single == again                                   # identical bits for identical inputs
        pc1 = np.array([pt[0] for pt in POINTS]); pc2 = np.array([pt[1] for pt in POINTS])
        al = np.array([pt[2] for pt in POINTS])
        assert eng.eval_batch(pc1, pc2, al).tolist() == single   # candidates in one pass: same bits
        # more candidates than fit in the kernel arguments -> parameters staged through HBM
        rep = 5
        big = eng.eval_batch(np.tile(pc1, (rep, 1)), np.tile(pc2, (rep, 1)), np.tile(al, rep))
        assert big.tolist() == single * rep


def test_both_many_evaluation_kernels_return_the_same_bits(sample10k, monkeypatch):
    """The common shapes (fp32 panel, NumPC 2 or 4, blobs that fit a stage) run llk_flow_kernel (parameters in the
    kernel arguments, eight warps per SM sub-partition); VB2_STREAM_KERNEL=queue sends them to llk_stream_kernel (task
    queue) like every other shape.  Same device functions, same order: the same bits, also across the launches a
    large batch is cut into and when jobs of several samples share a launch."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s4 = synth.make_sample(panel, n_pc=4, depth=25.0, alpha=0.04, seed=11, n_markers=6000)
    rng = np.random.default_rng(17)
    for prob, k in ((sample10k.problem, 2), (s4.problem, 4)):
        n = 333                                                   # more jobs than one launch's argument table holds
        pc1 = rng.normal(0, 0.02, (n, k)); pc2 = rng.normal(0, 0.02, (n, k)); al = rng.uniform(0.0, 1.0, n)
        al[:3] = (0.0, 1.0, 0.5)
        with vb.LLKEngine(prob) as eng, vb.LLKEngine(prob, shard_rank=1, shard_count=3, batched=True) as part:
            monkeypatch.delenv("VB2_STREAM_KERNEL", raising=False)
            flow = eng.eval_batch(pc1, pc2, al)
            flow_mixed = vb.eval_many([eng if j % 3 else part for j in range(n)], pc1, pc2, al)
            monkeypatch.setenv("VB2_STREAM_KERNEL", "queue")
            queue = eng.eval_batch(pc1, pc2, al)
            queue_mixed = vb.eval_many([eng if j % 3 else part for j in range(n)], pc1, pc2, al)
            monkeypatch.delenv("VB2_STREAM_KERNEL", raising=False)
            assert flow.tolist() == queue.tolist()
            assert flow_mixed.tolist() == queue_mixed.tolist()
            # what bench.py reports about a batch comes from the library, not from a constant
            assert eng.batch_plan(n) == {"kernel": "llk_flow_kernel", "kernel_launches": 3, "jobs_per_launch": 120}
            monkeypatch.setenv("VB2_STREAM_KERNEL", "queue")
            assert eng.batch_plan(n) == {"kernel": "llk_stream_kernel", "kernel_launches": 1, "jobs_per_launch": n}
            monkeypatch.delenv("VB2_STREAM_KERNEL", raising=False)
            for j in (0, 1, 2, 150, 332):
                assert flow[j] == eng.compute_mix_llks(pc1[j], pc2[j], al[j])


def test_launch_geometry_does_not_change_the_bits(sample10k, monkeypatch):
    """One evaluation per launch runs 4*kc warps per CTA (kc rounds of a bin in flight at once); whatever kc, the
    marginals of a bin are multiplied up in the same order, so the result is the same to the bit."""
    with vb.LLKEngine(sample10k.problem) as eng:
        want = [eng.compute_mix_llks(*pt) for pt in POINTS]
    for kc in ("1", "2", "3"):
        monkeypatch.setenv("VB2_LLK_LAT_KC", kc)
        with vb.LLKEngine(sample10k.problem) as eng:
            assert [eng.compute_mix_llks(*pt) for pt in POINTS] == want, kc


def test_session_resident_kernel_returns_the_same_bits(sample10k, monkeypatch):
    """vb2_llk_session_begin: one resident kernel, the sample in shared memory, evaluations through a doorbell."""
    import time
    monkeypatch.setenv("VB2_LLK_SESSION_IDLE_MS", "20")
    with vb.LLKEngine(sample10k.problem) as eng:
        want = [eng.compute_mix_llks(*pt) for pt in POINTS]
        eng.session_begin()
        eng.session_begin()                                            # idempotent
        assert [eng.compute_mix_llks(*pt) for pt in POINTS] == want
        eng.eval_begin(*POINTS[1])                                     # the split call rings the doorbell too
        assert eng.eval_end() == want[1]
        time.sleep(0.2)                                                # idle watchdog: the kernel has left by now ...
        assert [eng.compute_mix_llks(*pt) for pt in POINTS] == want    # ... and the next doorbell brings it back
        pc1 = np.array([pt[0] for pt in POINTS]); pc2 = np.array([pt[1] for pt in POINTS])
        al = np.array([pt[2] for pt in POINTS])
        assert eng.eval_batch(pc1, pc2, al).tolist() == want           # a batched call ends the session
        assert eng.compute_mix_llks(*POINTS[0]) == want[0]             # (launch per evaluation again)
        eng.session_begin()
        assert eng.compute_mix_llks(*POINTS[2]) == want[2]
        eng.session_end()
        eng.session_end()                                              # idempotent
        assert eng.compute_mix_llks(*POINTS[3]) == want[3]


def test_session_on_reference_fixture_and_unsupported_shapes(monkeypatch):
    with vb.LLKEngine(to_product(golden_problem(RESULT_PILEUP)), panel_dtype=vb.VB2_PANEL_FP64) as eng:
        eng.session_begin()
        for (pc1, pc2, a), want in zip(KAT_POINTS, KAT_RESULT):
            assert rel(eng.compute_mix_llks(pc1, pc2, a), want) <= REL_FP64
        eng.session_end()
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    deep = synth.make_sample(panel, n_pc=2, depth=200.0, alpha=0.02, seed=3, n_markers=600)
    with vb.LLKEngine(deep.problem) as eng:                            # blobs deeper than one stage: no session
        with pytest.raises(vb.VB2Error):
            eng.session_begin()
        assert np.isfinite(eng.compute_mix_llks(*POINTS[0]))


def test_stream_sync_wait_mode(sample10k):
    with vb.LLKEngine(sample10k.problem, spin=False) as a, vb.LLKEngine(sample10k.problem, spin=True) as b:
        for pt in POINTS[:3]:
            assert a.compute_mix_llks(*pt) == b.compute_mix_llks(*pt)


@pytest.mark.parametrize("n_shards", [2, 8])
def test_marker_shards_sum_to_whole(sample10k, n_shards):
    p = sample10k.problem
    with vb.LLKEngine(p) as whole:
        parts = [vb.LLKEngine(p, shard_rank=r, shard_count=n_shards) for r in range(n_shards)]
        try:
            assert sum(e.info()["reads_used"] for e in parts) == whole.info()["reads_used"]
            for pt in POINTS[:3]:
                total = sum(e.compute_mix_llks(*pt) for e in parts)
                assert rel(total, whole.compute_mix_llks(*pt)) <= 1e-12
        finally:
            for e in parts:
                e.close()


def test_eval_many_samples_one_launch():
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    samples = [synth.make_sample(panel, n_pc=2, depth=d, alpha=0.02, seed=100 + i, n_markers=3000 + 500 * i)
               for i, d in enumerate((8.0, 30.0, 55.0))]
    engines = [vb.LLKEngine(s.problem) for s in samples]
    try:
        order = [0, 1, 2, 1]                                    # a context may appear more than once
        pc1 = np.array([[0.01, 0.01], [0.02, -0.01], [0.0, 0.0], [0.03, 0.01]])
        pc2 = np.array([[0.01, 0.01], [-0.01, 0.027], [0.0, 0.0], [0.01, 0.01]])
        al = np.array([0.03, 0.05, 0.5, 0.2])
        got = vb.eval_many([engines[i] for i in order], pc1, pc2, al)
        want = [engines[i].compute_mix_llks(pc1[j], pc2[j], al[j]) for j, i in enumerate(order)]
        assert got.tolist() == want
    finally:
        for e in engines:
            e.close()


def test_eval_many_with_a_sample_that_has_no_usable_marker():
    """One empty (or fully filtered) pileup in a cohort answers 0.0 -- the reference's empty sum (h:231,:313) -- and
    does not disturb the other samples of the launch."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=20.0, alpha=0.02, seed=5, n_markers=2500)
    p = s.problem
    none = vb.PileupProblem(p.ud, p.means, np.full(p.n_marker, -1, np.int32), p.alt_base, np.zeros(1, np.int64),
                            np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    engines = [vb.LLKEngine(p, batched=True), vb.LLKEngine(none, batched=True), vb.LLKEngine(p, batched=True)]
    try:
        pcs = np.array([[0.01, 0.01], [0.0, 0.0], [0.02, -0.01]])
        al = np.array([0.03, 0.5, 0.1])
        got = vb.eval_many(engines, pcs, pcs, al)
        assert got[1] == 0.0
        assert got[0] == engines[0].compute_mix_llks(pcs[0], pcs[0], al[0])
        assert got[2] == engines[2].compute_mix_llks(pcs[2], pcs[2], al[2])
        assert vb.eval_many([engines[1]], pcs[1:2], pcs[1:2], al[1:2]).tolist() == [0.0]   # nothing to launch at all
    finally:
        for e in engines:
            e.close()


def test_batched_layout_and_many_jobs_per_launch(sample10k):
    """VB2_FLAG_BATCHED shards (fewer, deeper bins) sum to the whole sample; a launch may carry hundreds of jobs that
    reuse a few contexts (every job gets its own partial-sum slot)."""
    p = sample10k.problem
    ora = to_oracle(p)
    parts = [vb.LLKEngine(p, shard_rank=r, shard_count=4, batched=True) for r in range(4)]
    try:
        assert all(e.info()["grid_x"] < 40 for e in parts)             # 78 slices -> 3 CTAs' worth of bins, not 20
        n = 300
        rng = np.random.default_rng(5)
        pc1 = rng.normal(0, 0.02, (n, 2)); pc2 = rng.normal(0, 0.02, (n, 2)); al = rng.uniform(0.0, 0.5, n)
        total = np.zeros(n)
        for e in parts:
            total += vb.eval_many([e] * n, pc1, pc2, al)
        for j in (0, 1, 150, 299):
            assert rel(total[j], ora.compute_mix_llks(pc1[j], pc2[j], al[j])) <= REL_FP32
        mixed = vb.eval_many([parts[j % 4] for j in range(n)], pc1, pc2, al)   # contexts interleaved in one launch
        again = vb.eval_many([parts[j % 4] for j in range(n)], pc1, pc2, al)
        assert mixed.tolist() == again.tolist()
        assert mixed[5] == vb.eval_many([parts[1]], pc1[5:6], pc2[5:6], al[5:6])[0]
    finally:
        for e in parts:
            e.close()


def test_deep_coverage_multi_stage(monkeypatch):
    """200x depth with a tiny shared-memory stage: every slice needs several double-buffered TMA stages."""
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=200.0, alpha=0.02, seed=5, n_markers=2000)
    ora = to_oracle(s.problem)
    want = [ora.compute_mix_llks(*pt) for pt in POINTS[:3]]
    for stage in ("3", "7", "64"):
        monkeypatch.setenv("VB2_LLK_STAGE_WORDS", stage)
        with vb.LLKEngine(s.problem, panel_dtype=vb.VB2_PANEL_FP64) as eng:
            got = [eng.compute_mix_llks(*pt) for pt in POINTS[:3]]
        assert all(rel(g, w) <= REL_FP64 for g, w in zip(got, want)), (stage, got, want)


def test_known_af_mode(sample10k):
    p = sample10k.problem
    rng = np.random.default_rng(3)
    kaf = rng.uniform(0.0, 1.0, p.n_marker)
    kaf[:50] = 0.0                                               # exercises the 5e-5 clamp
    q = vb.PileupProblem(p.ud, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, kaf,
                         p.sanity_disabled, p.avg_depth, p.sd_depth)
    ora = to_oracle(q)
    with vb.LLKEngine(q) as eng:
        for a in (0.03, 0.3):
            assert rel(eng.compute_mix_llks([0, 0], [0, 0], a), ora.compute_mix_llks([0, 0], [0, 0], a)) <= REL_FP64


def test_edge_inputs():
    g = to_product(golden_problem(RESULT_PILEUP))
    # no usable marker -> the reference's empty sum, 0.0
    none = vb.PileupProblem(g.ud, g.means, np.full(g.n_marker, -1, np.int32), g.alt_base, np.zeros(1, np.int64),
                            np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    with vb.LLKEngine(none) as eng:
        assert eng.compute_mix_llks([0, 0], [0, 0], 0.5) == 0.0
        assert eng.eval_batch(np.zeros((3, 2)), np.zeros((3, 2)), np.full(3, 0.1)).tolist() == [0.0] * 3
    # qualities outside Phred+33 are clamped to [0, 93] (h:296-298); 'N' and unknown bases are class "other"
    ud = np.array([[0.5, -0.2], [0.1, 0.3]]); mu = np.array([0.8, 1.1])
    bases = np.frombuffer(b".,Aa" + b"NnGg*.", dtype=np.uint8)
    quals = np.array([0, 20, 33, 255, 40, 126, 33 + 93, 33 + 94, 60, 34], dtype=np.uint8)
    p = vb.PileupProblem(ud, mu, np.array([0, 1], np.int32), np.frombuffer(b"AG", dtype=np.uint8),
                         np.array([0, 4, 10], np.int64), bases, quals)
    ora = to_oracle(p)
    with vb.LLKEngine(p, panel_dtype=vb.VB2_PANEL_FP64) as eng:
        for a in (0.0, 0.2, 1.0):
            assert rel(eng.compute_mix_llks([0.1, 0.1], [-0.1, 0.2], a),
                       ora.compute_mix_llks([0.1, 0.1], [-0.1, 0.2], a)) <= REL_FP64
    # argument checking: wrong vector length is refused on the host side
    with vb.LLKEngine(g) as eng:
        with pytest.raises(ValueError):
            eng.compute_mix_llks([0.0], [0.0, 0.0], 0.1)


def test_full_size_invariants():
    """BASELINE config 2 (100k x 30x): size-independent properties at full size + oracle agreement."""
    panel = panels.load_bundled("1000g.phase3.100k.b37")
    s = synth.make_sample(panel, n_pc=2, depth=30.0, alpha=0.02, seed=1)
    p = s.problem
    ora = to_oracle(p, num_thread=os.cpu_count() or 1)
    with vb.LLKEngine(p) as eng:
        info = eng.info()
        assert info["reads_used"] == p.used_counts()[1] and info["markers_used"] == p.used_counts()[0]
        truth = (list(s.pc_contam), list(s.pc_intended), 0.02)
        l_true = eng.compute_mix_llks(*truth)
        assert rel(l_true, ora.compute_mix_llks(*truth)) <= REL_FP32
        # the likelihood prefers the generating parameters to the optimiser's start and to the null model
        assert l_true > eng.compute_mix_llks([0.01, 0.01], [0.01, 0.01], 0.03)
        assert l_true > eng.compute_mix_llks(truth[0], truth[1], 0.0)
        # additivity over marker shards (a checksum of checksums)
        parts = [vb.LLKEngine(p, shard_rank=r, shard_count=4) for r in range(4)]
        try:
            assert rel(sum(e.compute_mix_llks(*truth) for e in parts), l_true) <= 1e-12
        finally:
            for e in parts:
                e.close()


def test_timing_helpers_run_the_real_kernel(sample10k):
    """bench.py's C-side loops launch exactly what the evaluation calls launch."""
    engines = [vb.LLKEngine(sample10k.problem) for _ in range(3)]
    try:
        one = engines[:1]                    # (contexts of one timed loop must share a stream)
        assert vb.time_device(one, 2, 6, [0.01, 0.01], [0.01, 0.01], 0.03) > 0
        assert vb.time_device_many(engines, 1, 4, [0.01, 0.01], [0.01, 0.01], 0.03) > 0
        secs, last = vb.time_host(engines, 2, 6, [0.01, 0.01], [0.01, 0.01], 0.03)
        assert secs > 0
        # the last timed call evaluated pc1[0] = 0.01 + 1e-7 * 7 on engines[7 % 3]
        assert last == engines[1].compute_mix_llks([0.01 + 1e-7 * 7, 0.01], [0.01, 0.01], 0.03)
    finally:
        for e in engines:
            e.close()


def test_sharded_llk_single_rank_is_the_whole_sample(sample10k):
    """verifybamid_b200.distributed.ShardedLLK with world = 1 (no process group): same bits as the plain engine."""
    from verifybamid_b200.distributed import ShardedLLK
    s = ShardedLLK(sample10k.problem, device=0, rank=0, world=1)
    try:
        with vb.LLKEngine(sample10k.problem) as eng:
            for pt in POINTS[:3]:
                assert s.compute_mix_llks(*pt) == eng.compute_mix_llks(*pt)
    finally:
        s.close()


def test_abi_misuse_is_reported_not_crashed(sample10k):
    import ctypes
    lib = vb.load_library()
    with vb.LLKEngine(sample10k.problem) as eng:
        out = ctypes.c_double()
        assert lib.vb2_llk_eval_end(eng._ctx, ctypes.byref(out)) == 1          # nothing pending
        eng.eval_begin([0.01, 0.01], [0.01, 0.01], 0.03)
        with pytest.raises(vb.VB2Error):
            eng.eval_begin([0.01, 0.01], [0.01, 0.01], 0.03)                   # one evaluation per context at a time
        assert eng.eval_end() == eng.compute_mix_llks([0.01, 0.01], [0.01, 0.01], 0.03)
        with pytest.raises(vb.VB2Error):
            eng.eval_batch(np.zeros((0, 2)), np.zeros((0, 2)), np.zeros(0))    # empty batch
        assert b"pending" in lib.vb2_last_error(eng._ctx) or b"batch" in lib.vb2_last_error(eng._ctx)
    with pytest.raises(vb.VB2Error) as ei:
        vb.LLKEngine(sample10k.problem, device=99)
    assert ei.value.code == 2
    with pytest.raises(vb.VB2Error):
        vb.LLKEngine(sample10k.problem, shard_rank=3, shard_count=2)


def test_sixteen_pcs_and_multi_chunk_stage():
    """VB2_MAX_PC principal components: the fixed part of a blob grows, so even 30x reads need several stages."""
    rng = np.random.default_rng(12)
    panel = panels.load_bundled("1000g.phase3.10k.b37")
    s = synth.make_sample(panel, n_pc=4, depth=30.0, alpha=0.03, seed=2, n_markers=4000)
    p = s.problem
    ud16 = np.concatenate([p.ud, rng.normal(0, 1.0, (p.n_marker, 12))], axis=1)
    q = vb.PileupProblem(ud16, p.means, p.base_info_index, p.alt_base, p.info_offset, p.bases, p.quals, None,
                         p.sanity_disabled, p.avg_depth, p.sd_depth)
    ora = to_oracle(q)
    pc1 = rng.normal(0, 0.01, 16); pc2 = rng.normal(0, 0.01, 16)
    with vb.LLKEngine(q, panel_dtype=vb.VB2_PANEL_FP64) as eng:
        for a in (0.02, 0.4):
            assert rel(eng.compute_mix_llks(pc1, pc2, a), ora.compute_mix_llks(pc1, pc2, a)) <= REL_FP64
        two = eng.eval_batch(np.stack([pc1, pc2]), np.stack([pc2, pc1]), np.array([0.02, 0.4]))
        assert two[0] == eng.compute_mix_llks(pc1, pc2, 0.02) and two[1] == eng.compute_mix_llks(pc2, pc1, 0.4)
    with pytest.raises(vb.VB2Error):
        vb.LLKEngine(vb.PileupProblem(np.zeros((p.n_marker, 17)), p.means, p.base_info_index, p.alt_base, p.info_offset,
                                      p.bases, p.quals))


def test_runtime_layout_kernel_equals_specialised_kernel(sample10k, monkeypatch):
    """NumPC 2/4 with an fp32 panel runs a kernel instantiation with compile-time blob offsets; forcing the
    runtime-layout instantiation on the same sample must give the same bits."""
    with vb.LLKEngine(sample10k.problem) as eng:
        fast = [eng.compute_mix_llks(*pt) for pt in POINTS]
        fast_batch = eng.eval_batch(np.array([p[0] for p in POINTS]), np.array([p[1] for p in POINTS]),
                                    np.array([p[2] for p in POINTS])).tolist()
    monkeypatch.setenv("VB2_LLK_NO_SPEC", "1")
    with vb.LLKEngine(sample10k.problem) as eng:
        assert [eng.compute_mix_llks(*pt) for pt in POINTS] == fast
        assert eng.eval_batch(np.array([p[0] for p in POINTS]), np.array([p[1] for p in POINTS]),
                              np.array([p[2] for p in POINTS])).tolist() == fast_batch == fast


def test_minimize_on_the_device_matches_a_host_nelder_mead(sample10k):
    """vb2_llk_minimize against a plain Python transcription of AmoebaMinimizer::Minimize (MathGenMin.cpp:326-443) that
    calls the same engine one evaluation at a time: same evaluations, same best point, same bookkeeping."""
    import math
    p = sample10k.problem
    k = p.n_pc
    calls = []
    with vb.LLKEngine(p) as eng:
        def f(v):
            a = math.exp(v[2 * k]); a = a / (1.0 + a)
            val = 0.0 - eng.compute_mix_llks(v[:k], v[k:2 * k], a)
            calls.append((val, list(v[:k]), list(v[k:2 * k]), a))
            return val

        def amoeba(start, ftol, cycle_max=50000):
            dim = len(start); nv = dim + 1
            simplex = [[start[j] + (1.0 if j == i else 0.0) for j in range(dim)] for i in range(dim)] + [list(start)]
            y = [f(simplex[i]) for i in range(nv)]
            psum = list(simplex[0])
            for m in range(1, nv):
                for j in range(dim):
                    psum[j] += simplex[m][j]
            cycles = nv

            def trial(ihi, factor):
                fac = (1.0 - factor) / dim
                pt = [fac * psum[j] + (factor - fac) * simplex[ihi][j] for j in range(dim)]
                yt = f(pt)
                if yt < y[ihi]:
                    y[ihi] = yt
                    for j in range(dim):
                        psum[j] = psum[j] - simplex[ihi][j] + pt[j]
                    simplex[ihi] = pt
                return yt
            while True:
                if y[0] > y[1]:
                    ilo = inhi = 1; ihi = 0
                else:
                    ilo = inhi = 0; ihi = 1
                for i in range(2, nv):
                    if y[i] <= y[ilo]:
                        ilo = i
                    elif y[i] > y[ihi]:
                        inhi = ihi; ihi = i
                    elif y[i] > y[inhi]:
                        inhi = i
                rtol = 2 * abs(y[ihi] - y[ilo]) / (abs(y[ihi]) + abs(y[ilo]) + 3.0e-10)
                if rtol < ftol:
                    return simplex[ilo], y[ilo], cycles
                assert cycles <= cycle_max
                cycles += 2
                yt = trial(ihi, -1.0)
                if yt <= y[ilo]:
                    trial(ihi, 2.0)
                elif yt >= y[inhi]:
                    ysave = y[ihi]
                    yt = trial(ihi, 0.5)
                    if yt >= ysave:
                        for i in range(nv):
                            if i != ilo:
                                simplex[i] = [(simplex[i][j] + simplex[ilo][j]) * 0.5 for j in range(dim)]
                                y[i] = f(simplex[i])
                        cycles += dim
                        psum = list(simplex[0])
                        for m in range(1, nv):
                            for j in range(dim):
                                psum[j] += simplex[m][j]
                else:
                    cycles -= 1

        start = [0.01] * (2 * k) + [math.log(0.03 / 0.97)]
        eng.session_begin()
        want_pt, want_f, want_cycles = amoeba(start, 1e-8)
        got = eng.minimize(start, list(range(k)), list(range(k, 2 * k)), 2 * k, ftol=1e-8)
        # a second search in the same session, a model with fixed parts: alpha fixed, one set of PCs for both samples
        got2 = eng.minimize([0.01] * k, list(range(k)), list(range(k)), -1, alpha_fixed=0.05, ftol=1e-8, llk1=got["llk1"])
        one = eng.compute_mix_llks(got2["point"], got2["point"], 0.05)       # evaluations still work after a search
        eng.session_end()
    assert got["converged"] and got["evals"] == len(calls), (got["evals"], len(calls))
    assert got["cycle_count"] == want_cycles
    assert got["point"] == want_pt                                        # the same trajectory, bit for bit
    # (alpha = InvLogit(v) goes through exp(): the device's and the host's may differ in the last bit)
    assert rel(got["fmin"], want_f) <= 1e-13
    best = min(calls, key=lambda c: c[0])
    assert got["improved"] and rel(got["llk1"], best[0]) <= 1e-13
    assert got["best_pc_contam"] == best[1] and got["best_pc_intended"] == best[2] and rel(got["best_alpha"], best[3]) <= 1e-15
    assert got2["converged"] and got2["fmin"] == 0.0 - one
    assert got2["fmin"] >= got["fmin"] - 1e-9 * abs(got["fmin"])          # a constrained model cannot beat the free one


@pytest.mark.skipif(vb.device_count() < 2, reason="needs two GPUs")
def test_peer_stores_give_the_same_sums_as_nccl():
    """vb2_peer_*: the shard sums pushed over NVLink by the reduce kernel + gather kernel, against an NCCL all-reduce of the
    same shard sums, two processes (tools/collective_ab.py under torchrun)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cp = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                         "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "collective_ab.py")],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert cp.returncode == 0, cp.stdout[-2000:]
    assert "same sums: True" in cp.stdout, cp.stdout[-2000:]
