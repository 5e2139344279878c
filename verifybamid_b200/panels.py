"""Reference SVD panels: load the packed copies that ship in
verifybamid_b200/data/ and expand them back to the plain-text <prefix>.UD/.mu/.bed files the
CLI (and the reference) read with --SVDPrefix.

The packed files are made by tools/make_panel_npz.py from the reference's resource/ directory;
values are the float64 that `operator>>` parses, so text written here re-parses to the same
doubles (repr round-trip).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

BUNDLED = {
    "1000g.phase3.100k.b37": "1000g.phase3.100k.b37.npz",
    "1000g.phase3.10k.b37": "1000g.phase3.10k.b37.npz",
    "hgdp.100k.b37": "hgdp.100k.b37.npz",
}


@dataclass
class PanelData:
    name: str
    ud: np.ndarray        # [M, 4] float64
    mu: np.ndarray        # [M] float64 (mean genotype in [0,2])
    chrom: List[str]
    pos: np.ndarray       # [M] int64, 1-based
    ref: List[str]        # REF column (string; the reference reads its first char)
    alt: List[str]        # ALT column (string; the reference reads its first char)
    v: np.ndarray         # [N, 4] float64: sample PCs (.V), used only to pick realistic PCs

    @property
    def n_marker(self) -> int:
        return int(self.ud.shape[0])

    def alt_char(self) -> np.ndarray:
        """ALT as the reference sees it: ONE char (ContaminationEstimator.cpp:417,429)."""
        return np.frombuffer("".join(a[0] for a in self.alt).encode(), dtype=np.uint8).copy()

    def ref_char(self) -> np.ndarray:
        return np.frombuffer("".join(r[0] for r in self.ref).encode(), dtype=np.uint8).copy()


def load_bundled(name: str) -> PanelData:
    if name not in BUNDLED:
        raise KeyError("unknown bundled panel %r (have: %s)" % (name, ", ".join(sorted(BUNDLED))))
    z = np.load(os.path.join(DATA_DIR, BUNDLED[name]))
    return PanelData(name, z["ud"], z["mu"], [str(c) for c in z["chrom"]], z["pos"],
                     [str(r) for r in z["ref"]], [str(a) for a in z["alt"]], z["v"])


def write_text_panel(panel: PanelData, prefix: str) -> str:
    """Write <prefix>.UD/.mu/.bed in the reference's on-disk format; returns prefix."""
    with open(prefix + ".UD", "w") as f:
        for row in panel.ud:
            f.write("\t".join(repr(float(x)) for x in row) + "\n")
    with open(prefix + ".mu", "w") as f:
        for c, p, r, a, m in zip(panel.chrom, panel.pos, panel.ref, panel.alt, panel.mu):
            f.write("%s:%d_%s/%s\t%r\n" % (c, int(p), r, a, float(m)))
    with open(prefix + ".bed", "w") as f:
        for c, p, r, a in zip(panel.chrom, panel.pos, panel.ref, panel.alt):
            f.write("%s\t%d\t%d\t%s\t%s\n" % (c, int(p) - 1, int(p), r, a))
    return prefix
