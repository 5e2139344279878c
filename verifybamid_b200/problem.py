"""Flat (CSR) image of what the reference's ComputeMixLLKs reads -- the argument of the C ABI.

Field-for-field the same as `vb2_llk_desc` in include/vb2_llk.h (which cites the reference
object each field replaces), held as numpy arrays.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass
class PileupProblem:
    ud: np.ndarray                 # [M, k] float64       ContaminationEstimator::UD
    means: np.ndarray              # [M] float64          ContaminationEstimator::means
    base_info_index: np.ndarray    # [M] int32, -1 absent resolvedMarkers[i].baseInfoIndex
    alt_base: np.ndarray           # [M] uint8            resolvedMarkers[i].altBase
    info_offset: np.ndarray        # [n_info+1] int64     CSR over viewer.baseInfo
    bases: np.ndarray              # [R] uint8            viewer.baseInfo chars
    quals: np.ndarray              # [R] uint8            viewer.qualInfo chars (Phred+33)
    known_af: Optional[np.ndarray] = None   # [M] float64 resolvedMarkers[i].knownAFValue
    sanity_disabled: bool = True
    avg_depth: float = 0.0
    sd_depth: float = 0.0
    n_marker_total: int = 0

    def __post_init__(self):
        self.ud = np.ascontiguousarray(self.ud, dtype=np.float64)
        if self.ud.ndim != 2:
            raise ValueError("ud must be [M, k]")
        self.means = np.ascontiguousarray(self.means, dtype=np.float64)
        self.base_info_index = np.ascontiguousarray(self.base_info_index, dtype=np.int32)
        self.alt_base = np.ascontiguousarray(self.alt_base, dtype=np.uint8)
        self.info_offset = np.ascontiguousarray(self.info_offset, dtype=np.int64)
        self.bases = np.ascontiguousarray(self.bases, dtype=np.uint8)
        self.quals = np.ascontiguousarray(self.quals, dtype=np.uint8)
        if self.known_af is not None:
            self.known_af = np.ascontiguousarray(self.known_af, dtype=np.float64)
        m = self.ud.shape[0]
        if not (len(self.means) == len(self.base_info_index) == len(self.alt_base) == m):
            raise ValueError("per-marker arrays differ in length")
        if len(self.bases) != len(self.quals):
            raise ValueError("bases and quals differ in length")
        if not self.n_marker_total:
            self.n_marker_total = int(m)

    @property
    def n_marker(self) -> int:
        return int(self.ud.shape[0])

    @property
    def n_pc(self) -> int:
        return int(self.ud.shape[1])

    def depths(self) -> np.ndarray:
        """Per panel marker depth (0 for absent markers)."""
        d = np.zeros(self.n_marker, dtype=np.int64)
        has = self.base_info_index >= 0
        idx = self.base_info_index[has]
        d[has] = self.info_offset[idx + 1] - self.info_offset[idx]
        return d

    def used_mask(self) -> np.ndarray:
        """Marker skip rules of ComputeMixLLKs (ContaminationEstimator.h:238-249)."""
        d = self.depths()
        keep = (self.base_info_index >= 0) & (d > 0)
        if not self.sanity_disabled:
            lo = self.avg_depth - 3 * self.sd_depth
            hi = self.avg_depth + 3 * self.sd_depth
            keep &= ~((d < lo) | (d > hi))
        return keep

    def used_counts(self) -> Tuple[int, int]:
        keep = self.used_mask()
        return int(keep.sum()), int(self.depths()[keep].sum())

    def subset_markers(self, rows: np.ndarray) -> "PileupProblem":
        """A problem restricted to the given panel rows (pileup arrays are shared)."""
        rows = np.asarray(rows)
        return PileupProblem(self.ud[rows], self.means[rows], self.base_info_index[rows], self.alt_base[rows],
                             self.info_offset, self.bases, self.quals,
                             None if self.known_af is None else self.known_af[rows],
                             self.sanity_disabled, self.avg_depth, self.sd_depth, len(rows))
