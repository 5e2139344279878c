"""oracle/vb2_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python side of the CPU oracle: restates the reference's *host* data path for the
contamination-LLK hot path (file readers, text-pileup parser, marker resolution, depth
sanity filter) in plain Python/numpy, and binds oracle/liboracle.so (llk_oracle.c, the C
restatement of ComputeMixLLKs + AmoebaMinimizer + OptimizeLLK) through ctypes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  Parity status: PINNED (see the header of llk_oracle.c and
tests/test_oracle.py).  All file:line citations are relative to the reference tree.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "vb2_ref")
MAXDIM = 64


# --------------------------------------------------------------------------------------------
# build helpers
# --------------------------------------------------------------------------------------------
def build(quiet: bool = True) -> None:
    """Compile liboracle.so and (when /root/reference is present) oracle/_ref/vb2_ref."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def ref_available() -> bool:
    return os.access(REF_BIN, os.X_OK)


# --------------------------------------------------------------------------------------------
# panel readers: ContaminationEstimator.cpp:342-373 (ReadMatrixUD), :440-459 (ReadMean),
# :413-438 (ReadChooseBed)
# --------------------------------------------------------------------------------------------
@dataclass
class Panel:
    ud: np.ndarray            # [M, k] float64, first k columns of .UD
    mu: np.ndarray            # [M] float64, column 2 of .mu
    chrom: List[str]          # PosVec[i].first
    pos: np.ndarray           # [M] int64, PosVec[i].second (1-based)
    ref: List[str]
    alt: List[str]            # first char of the ALT column ("G,T" -> "G")
    choose_bed: Dict[str, Dict[int, Tuple[str, str]]] = field(default_factory=dict)

    @property
    def n_marker(self) -> int:
        return int(self.ud.shape[0])


def read_panel(prefix: str, n_pc: int) -> Panel:
    ud_rows = []
    with open(prefix + ".UD") as f:
        for line in f:
            tok = line.split()
            if len(tok) < n_pc:
                raise ValueError("--NumPC exceeds the PCs in the .UD file")  # cpp:358-363
            ud_rows.append([float(t) for t in tok[:n_pc]])
    ud = np.asarray(ud_rows, dtype=np.float64).reshape(len(ud_rows), n_pc)
    mu = []
    with open(prefix + ".mu") as f:
        for line in f:
            tok = line.split()
            mu.append(float(tok[1]))
    chrom, pos, ref, alt = [], [], [], []
    choose: Dict[str, Dict[int, Tuple[str, str]]] = {}
    with open(prefix + ".bed") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            c, p = tok[0], int(tok[2])        # ss >> chr >> pos >> pos  (cpp:428)
            r, a = tok[3][0], tok[4][0]       # ss >> ref >> alt as single chars (cpp:429)
            chrom.append(c); pos.append(p); ref.append(r); alt.append(a)
            choose.setdefault(c, {})[p] = (r, a)
    return Panel(ud, np.asarray(mu, dtype=np.float64), chrom, np.asarray(pos, dtype=np.int64), ref, alt, choose)


# --------------------------------------------------------------------------------------------
# text pileup: SimplePileupViewer.cpp:711-746 (ParsePileupSeqBasesOnly), :748-833 (ReadPileup)
# --------------------------------------------------------------------------------------------
_KEEP = set(b"ACGTNacgtn")


def parse_pileup_seq_bases_only(seq: bytes, qual: bytes) -> Tuple[bytes, bytes]:
    pseq, pqual = bytearray(), bytearray()
    i, iq, n = 0, 0, len(seq)
    while i < n:
        c = seq[i]
        if c in (0x2B, 0x2D):                      # '+' / '-': skip digits and that many chars
            t = i + 1
            while t < n and 0x30 <= seq[t] <= 0x39:
                t += 1
            digit_len = t - (i + 1)
            clip = int(seq[i + 1:t])               # std::stoi: throws on no digits, as here
            i += digit_len + clip
        elif c == 0x5E:                            # '^': skip the mapping-quality char
            i += 1
        elif c in (0x2E, 0x2C) or c in _KEEP:      # '.' ',' ACGTN acgtn: keep, consume a qual
            pseq.append(c)
            pqual.append(qual[iq])
            iq += 1
        elif c in (0x2A, 0x23):                    # '*' '#': dropped but consume a qual
            iq += 1
        i += 1
    return bytes(pseq), bytes(pqual)


@dataclass
class Viewer:
    base_info: List[bytes]
    qual_info: List[bytes]
    pos_index: Dict[str, Dict[int, int]]
    num_bases: int
    effective_num_site: int
    avg_depth: float
    sd_depth: float = 0.0


def read_pileup(path: str, choose_bed: Dict[str, Dict[int, Tuple[str, str]]]) -> Viewer:
    base_info: List[bytes] = []
    qual_info: List[bytes] = []
    pos_index: Dict[str, Dict[int, int]] = {}
    num_bases = 0
    eff = 0
    p_chr, p_pos, ref_allele, seq, qual = b"", 0, b"", b"", b""
    with open(path, "rb") as f:
        for line in f:
            tok = line.split()
            # ss >> pChr >> pPos >> refAllele >> depth >> seq >> qual: a short line leaves the
            # remaining variables at their previous values (seq/qual were reset to "" below).
            if len(tok) > 0: p_chr = tok[0]
            if len(tok) > 1: p_pos = int(tok[1])
            if len(tok) > 2: ref_allele = tok[2]
            if len(tok) > 4: seq = tok[4]
            if len(tok) > 5: qual = tok[5]
            if (b"." in seq or b"," in seq) and ref_allele == b".":
                raise ValueError("Pileup format error: cannot find ref allele")       # cpp:771-778
            pseq, pqual = parse_pileup_seq_bases_only(seq, qual)
            seq, qual = pseq, pqual
            depth = len(pqual)
            c = p_chr.decode()
            if c not in choose_bed or p_pos not in choose_bed[c]:
                continue                                                                # cpp:787-790
            existed = c in pos_index and p_pos in pos_index[c]
            if not existed:
                pos_index.setdefault(c, {})[p_pos] = len(base_info)
                base_info.append(seq)
                qual_info.append(qual)
            # duplicates: the merged copy is discarded (cpp:812-824) but the counters still grow
            num_bases += depth
            seq, qual = b"", b""
            eff += 1
    avg = (num_bases / eff) if eff else float("nan")                                   # cpp:831
    return Viewer(base_info, qual_info, pos_index, num_bases, eff, avg)


# --------------------------------------------------------------------------------------------
# BuildResolvedMarkers (ContaminationEstimator.cpp:67-86), IsSanityCheckOK (:543-587)
# --------------------------------------------------------------------------------------------
def build_resolved_markers(panel: Panel, viewer: Viewer,
                           known_af: Optional[Dict[str, Dict[int, float]]] = None):
    m = panel.n_marker
    idx = np.full(m, -1, dtype=np.int32)
    alt = np.zeros(m, dtype=np.uint8)
    kaf = np.zeros(m, dtype=np.float64)
    for i in range(m):
        c, p = panel.chrom[i], int(panel.pos[i])
        row = viewer.pos_index.get(c)
        if row is None or p not in row:
            continue
        idx[i] = row[p]
        alt[i] = ord(panel.choose_bed[c][p][1])
        if known_af is not None:
            kaf[i] = known_af.get(c, {}).get(p, 0.0)
    return idx, alt, (kaf if known_af is not None else None)


def sanity_check(panel: Panel, viewer: Viewer) -> bool:
    """Mutates viewer.sd_depth / effective_num_site exactly like IsSanityCheckOK."""
    n = panel.n_marker
    acc = viewer.sd_depth
    depths = []
    for i in range(n):
        c, p = panel.chrom[i], int(panel.pos[i])
        row = viewer.pos_index.get(c)
        if row is None or p not in row:
            continue
        d = len(viewer.base_info[row[p]])
        depths.append(d)
        acc += d * d
    viewer.sd_depth = math.sqrt(acc / viewer.effective_num_site - viewer.avg_depth * viewer.avg_depth)
    lo = viewer.avg_depth - 3 * viewer.sd_depth
    hi = viewer.avg_depth + 3 * viewer.sd_depth
    viewer.effective_num_site = sum(1 for d in depths if not (d == 0 or d < lo or d > hi))
    return viewer.effective_num_site > 1000 and viewer.effective_num_site > n * 0.1


_NUM = __import__("re").compile(r"\s*([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?))")


def read_known_af(path: str) -> Dict[str, Dict[int, float]]:
    """ContaminationEstimator.cpp:461-487: `ss >> chr >> pos >> pos >> ref >> alt >> AF` with ref/alt single
    chars.  Stream semantics matter: after a multi-allelic ALT ("A,G") the extraction of AF starts at ",G",
    fails, and (C++11) stores 0 -- the reference really uses AF = 0 for those rows."""
    out: Dict[str, Dict[int, float]] = {}
    with open(path) as f:
        for line in f:
            tok = line.split(None, 3)
            if len(tok) < 4:
                continue
            rest = tok[3].lstrip()[1:].lstrip()[1:]        # drop the ref char, then the alt char
            m = _NUM.match(rest)
            out.setdefault(tok[0], {})[int(tok[2])] = float(m.group(1)) if m else 0.0
    return out


# --------------------------------------------------------------------------------------------
# flat problem + ctypes binding of llk_oracle.c
# --------------------------------------------------------------------------------------------
class _CProblem(ctypes.Structure):
    _fields_ = [("n_marker", ctypes.c_int), ("n_pc", ctypes.c_int),
                ("ud", ctypes.c_void_p), ("ud_stride", ctypes.c_int),
                ("means", ctypes.c_void_p), ("base_info_index", ctypes.c_void_p),
                ("alt_base", ctypes.c_void_p), ("known_af", ctypes.c_void_p),
                ("info_offset", ctypes.c_void_p), ("bases", ctypes.c_void_p), ("quals", ctypes.c_void_p),
                ("sanity_disabled", ctypes.c_int), ("avg_depth", ctypes.c_double),
                ("sd_depth", ctypes.c_double), ("num_thread", ctypes.c_int)]


class _CModel(ctypes.Structure):
    _fields_ = [("is_heter", ctypes.c_int), ("is_pc_fixed", ctypes.c_int), ("is_alpha_fixed", ctypes.c_int),
                ("alpha", ctypes.c_double), ("pc_fixed", ctypes.c_double * MAXDIM), ("epsilon", ctypes.c_double),
                ("global_pc", ctypes.c_double * MAXDIM), ("global_pc2", ctypes.c_double * MAXDIM),
                ("global_alpha", ctypes.c_double), ("llk1", ctypes.c_double), ("llk0", ctypes.c_double),
                ("evals", ctypes.c_longlong)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vb2o_compute_mix_llks.restype = ctypes.c_double
        _lib.vb2o_compute_mix_llks.argtypes = [ctypes.POINTER(_CProblem), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        _lib.vb2o_optimize_llk.restype = ctypes.c_int
        _lib.vb2o_optimize_llk.argtypes = [ctypes.POINTER(_CProblem), ctypes.POINTER(_CModel)]
        _lib.vb2o_used_counts.restype = None
        _lib.vb2o_used_counts.argtypes = [ctypes.POINTER(_CProblem), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]
    return _lib


@dataclass
class Problem:
    """Flat (CSR) image of what ComputeMixLLKs reads: UD, means, resolvedMarkers, viewer data."""
    ud: np.ndarray                 # [M, k] float64
    means: np.ndarray              # [M] float64
    base_info_index: np.ndarray    # [M] int32 (-1 absent)
    alt_base: np.ndarray           # [M] uint8
    info_offset: np.ndarray        # [n_info+1] int64
    bases: np.ndarray              # [R] uint8 (pileup base chars)
    quals: np.ndarray              # [R] uint8 (qual chars, Phred+33)
    known_af: Optional[np.ndarray] = None
    sanity_disabled: bool = True
    avg_depth: float = 0.0
    sd_depth: float = 0.0
    num_thread: int = 1
    n_marker_total: int = 0        # NumMarker as printed in .selfSM (#SNPS)

    def __post_init__(self):
        self.ud = np.ascontiguousarray(self.ud, dtype=np.float64)
        self.means = np.ascontiguousarray(self.means, dtype=np.float64)
        self.base_info_index = np.ascontiguousarray(self.base_info_index, dtype=np.int32)
        self.alt_base = np.ascontiguousarray(self.alt_base, dtype=np.uint8)
        self.info_offset = np.ascontiguousarray(self.info_offset, dtype=np.int64)
        self.bases = np.ascontiguousarray(self.bases, dtype=np.uint8)
        self.quals = np.ascontiguousarray(self.quals, dtype=np.uint8)
        if self.known_af is not None:
            self.known_af = np.ascontiguousarray(self.known_af, dtype=np.float64)
        if not self.n_marker_total:
            self.n_marker_total = int(self.ud.shape[0])

    @property
    def n_pc(self) -> int:
        return int(self.ud.shape[1])

    def c_struct(self) -> _CProblem:
        p = _CProblem()
        p.n_marker = int(self.ud.shape[0]); p.n_pc = self.n_pc
        p.ud = self.ud.ctypes.data; p.ud_stride = self.n_pc
        p.means = self.means.ctypes.data
        p.base_info_index = self.base_info_index.ctypes.data
        p.alt_base = self.alt_base.ctypes.data
        p.known_af = self.known_af.ctypes.data if self.known_af is not None else None
        p.info_offset = self.info_offset.ctypes.data
        p.bases = self.bases.ctypes.data; p.quals = self.quals.ctypes.data
        p.sanity_disabled = int(self.sanity_disabled)
        p.avg_depth = float(self.avg_depth); p.sd_depth = float(self.sd_depth)
        p.num_thread = int(self.num_thread)
        return p

    # ---- the oracle proper ---------------------------------------------------------------
    def compute_mix_llks(self, pc_contam: Sequence[float], pc_intended: Sequence[float], alpha: float) -> float:
        a = np.ascontiguousarray(pc_contam, dtype=np.float64)
        b = np.ascontiguousarray(pc_intended, dtype=np.float64)
        assert a.size == self.n_pc and b.size == self.n_pc
        cp = self.c_struct()
        return float(lib().vb2o_compute_mix_llks(ctypes.byref(cp), a.ctypes.data, b.ctypes.data, float(alpha)))

    def used_counts(self) -> Tuple[int, int]:
        cp = self.c_struct()
        m, r = ctypes.c_longlong(), ctypes.c_longlong()
        lib().vb2o_used_counts(ctypes.byref(cp), ctypes.byref(m), ctypes.byref(r))
        return int(m.value), int(r.value)

    def optimize(self, within_ancestry: bool = False, fix_pc: Optional[Sequence[float]] = None,
                 fix_alpha: Optional[float] = None, epsilon: float = 1e-8) -> dict:
        m = _CModel()
        m.is_heter = int(not within_ancestry)
        m.alpha = 0.5
        m.epsilon = epsilon
        if fix_pc is not None:                      # main.cpp:291-307
            for i in range(self.n_pc):
                m.pc_fixed[i] = float(fix_pc[i])
            m.is_pc_fixed = 1
        elif fix_alpha is not None:                 # main.cpp:308-312
            m.alpha = float(fix_alpha)
            m.is_alpha_fixed = 1
        if self.known_af is not None:               # main.cpp:313-318
            m.is_pc_fixed = 1
            m.is_heter = 0
        cp = self.c_struct()
        rc = lib().vb2o_optimize_llk(ctypes.byref(cp), ctypes.byref(m))
        if rc != 0:
            raise RuntimeError("vb2o_optimize_llk failed")
        k = self.n_pc
        return {"alpha": m.global_alpha, "llk1": m.llk1, "llk0": m.llk0,
                "pc_contam": [m.global_pc[i] for i in range(k)],
                "pc_intended": [m.global_pc2[i] for i in range(k)],
                "evals": int(m.evals)}


def problem_from_files(svd_prefix: str, pileup: str, n_pc: int = 2, disable_sanity: bool = False,
                       known_af_path: Optional[str] = None, num_thread: int = 1) -> Problem:
    """main.cpp:283-333 + :371-379 for --PileupFile input, producing the flat Problem."""
    panel = read_panel(svd_prefix, n_pc)
    viewer = read_pileup(pileup, panel.choose_bed)
    kaf = read_known_af(known_af_path) if known_af_path else None
    if not disable_sanity:
        if not sanity_check(panel, viewer):
            raise RuntimeError("Insufficient Available markers")
    idx, alt, kaf_arr = build_resolved_markers(panel, viewer, kaf)
    lens = np.fromiter((len(b) for b in viewer.base_info), dtype=np.int64, count=len(viewer.base_info))
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(b"".join(viewer.base_info), dtype=np.uint8)
    quals = np.frombuffer(b"".join(viewer.qual_info), dtype=np.uint8)
    return Problem(panel.ud, panel.mu, idx, alt, off, bases, quals, kaf_arr, disable_sanity,
                   viewer.avg_depth, viewer.sd_depth, num_thread, panel.n_marker)


# --------------------------------------------------------------------------------------------
# output formatting: default ostream precision (6 significant digits, %g-like)
# ContaminationEstimator.cpp:176-180 (.Ancestry), main.cpp:386-411 (.selfSM)
# --------------------------------------------------------------------------------------------
def ostream_double(x: float) -> str:
    return "%g" % x


def format_ancestry(pc_contam: Sequence[float], pc_intended: Sequence[float]) -> str:
    out = ["PC\tContaminatingSample\tIntendedSample"]
    for i, (a, b) in enumerate(zip(pc_contam, pc_intended)):
        out.append("%d\t%s\t%s" % (i + 1, ostream_double(a), ostream_double(b)))
    return "\n".join(out) + "\n"


# --------------------------------------------------------------------------------------------
# oracle/_ref runner (the reference's own sources compiled here)
# --------------------------------------------------------------------------------------------
def run_ref(args: Sequence[str], cwd: Optional[str] = None, timeout: float = 3600.0) -> List[dict]:
    """Run oracle/_ref/vb2_ref and return its machine-readable VB2REF records."""
    if not ref_available():
        raise FileNotFoundError(REF_BIN)
    cp = subprocess.run([REF_BIN, *args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                        timeout=timeout, check=True, text=True)
    return [json.loads(l[len("VB2REF "):]) for l in cp.stdout.splitlines() if l.startswith("VB2REF ")]
