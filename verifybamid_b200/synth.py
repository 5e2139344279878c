"""Synthetic contaminated pileups on a real SVD panel (BASELINE.json configs).

One generator, one seed -> both views of the same sample:
  * `PileupProblem`  (flat arrays, what the C ABI / GPU engine takes), and
  * samtools-pileup text (what `--PileupFile` of the CLI and of the reference binary read).

Model, per marker i of the panel: AF_i = clamp((UD_i . PC + mu_i)/2, 5e-5, 1-5e-5) for the intended
and the contaminating sample; genotypes ~ Binomial(2, AF); depth ~ Poisson(depth); each read
comes from the contaminant with probability alpha; allele = ALT w.p. g/2; q ~ U{q_lo..q_hi}; with
probability 10^(-q/10) the base is replaced by one of the other three; strand 50/50 ('.'/',' for
REF, upper/lower-case letter otherwise); qual char = q + 33.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .panels import PanelData
from .problem import PileupProblem

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class SyntheticSample:
    problem: PileupProblem
    panel: PanelData
    depth: np.ndarray          # [M] reads per marker
    pc_intended: np.ndarray    # [k] truth
    pc_contam: np.ndarray      # [k] truth
    alpha: float               # truth
    seed: int

    def write_pileup(self, path: str) -> str:
        """samtools-pileup text: chr pos ref depth seq qual (one line per panel marker)."""
        p, pan = self.problem, self.panel
        off = p.info_offset
        bases = p.bases.tobytes()
        quals = p.quals.tobytes()
        refc = pan.ref_char()
        with open(path, "wb") as f:
            for i in range(p.n_marker):
                b = int(p.base_info_index[i])
                lo, hi = int(off[b]), int(off[b + 1])
                if hi > lo:
                    f.write(b"%s\t%d\t%c\t%d\t%s\t%s\n" % (pan.chrom[i].encode(), int(pan.pos[i]), int(refc[i]),
                                                           hi - lo, bases[lo:hi], quals[lo:hi]))
                else:  # samtools prints '*' for an empty column
                    f.write(b"%s\t%d\t%c\t0\t*\t*\n" % (pan.chrom[i].encode(), int(pan.pos[i]), int(refc[i])))
        return path


def make_sample(panel: PanelData, n_pc: int = 2, depth: float = 30.0, alpha: float = 0.02, seed: int = 1,
                sanity_check: bool = True, q_lo: int = 20, q_hi: int = 40,
                n_markers: Optional[int] = None) -> SyntheticSample:
    """Generate one synthetic sample on `panel` (first `n_markers` rows if given)."""
    rng = np.random.default_rng(seed)
    m = panel.n_marker if n_markers is None else min(int(n_markers), panel.n_marker)
    ud = panel.ud[:m, :n_pc]
    mu = panel.mu[:m]
    pc_int = panel.v[0, :n_pc].copy()
    pc_con = panel.v[panel.v.shape[0] // 2, :n_pc].copy()
    af_int = np.clip((ud @ pc_int + mu) / 2.0, 5e-5, 1 - 5e-5)
    af_con = np.clip((ud @ pc_con + mu) / 2.0, 5e-5, 1 - 5e-5)
    g_int = rng.binomial(2, af_int)
    g_con = rng.binomial(2, af_con)
    dep = rng.poisson(depth, m).astype(np.int64)
    r = int(dep.sum())
    marker = np.repeat(np.arange(m), dep)
    from_con = rng.random(r) < alpha
    g = np.where(from_con, g_con[marker], g_int[marker])
    is_alt = rng.random(r) < g / 2.0
    q = rng.integers(q_lo, q_hi + 1, r)
    err = rng.random(r) < np.power(10.0, -q / 10.0)
    refc = panel.ref_char()[:m]
    altc = panel.alt_char()[:m]
    true_base = np.where(is_alt, altc[marker], refc[marker]).astype(np.uint8)
    idx = np.zeros(r, dtype=np.int64)
    for j, c in enumerate(_ACGT):
        idx[(true_base & 0xDF) == c] = j
    shift = rng.integers(1, 4, r)
    obs = np.where(err, _ACGT[(idx + shift) % 4], true_base & 0xDF).astype(np.uint8)
    fwd = rng.random(r) < 0.5
    is_ref = obs == (refc[marker] & 0xDF)
    chars = np.where(is_ref, np.where(fwd, ord("."), ord(",")), np.where(fwd, obs, obs | 0x20)).astype(np.uint8)
    quals = (q + 33).astype(np.uint8)

    off = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(dep, out=off[1:])
    # every panel marker has a pileup line -> viewer.baseInfo entry i belongs to marker i
    # (SimplePileupViewer.cpp:748-833); avgDepth = numBases / effectiveNumSite (:831)
    avg = r / m
    sd = 0.0
    if sanity_check:
        # ContaminationEstimator.cpp:553-565 (IsSanityCheckOK): sd from sum of squared depths
        sd = float(np.sqrt(float((dep * dep).sum()) / m - avg * avg))
    prob = PileupProblem(ud, mu, np.arange(m, dtype=np.int32), altc, off, chars, quals, None,
                         sanity_disabled=not sanity_check, avg_depth=avg, sd_depth=sd, n_marker_total=m)
    sub = panel if m == panel.n_marker else PanelData(panel.name, panel.ud[:m], panel.mu[:m], panel.chrom[:m],
                                                      panel.pos[:m], panel.ref[:m], panel.alt[:m], panel.v)
    return SyntheticSample(prob, sub, dep, pc_int, pc_con, alpha, seed)
