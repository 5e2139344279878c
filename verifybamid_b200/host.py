"""ctypes binding of libvb2host.so: the engine's C++ HOST side (panel readers, text-pileup parser, marker
resolution, depth sanity filter, Nelder-Mead) callable from Python without a GPU.

`load_problem()` runs exactly what the CLI runs before it hands the sample to the GPU and returns the flat
`PileupProblem` the C ABI takes.
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

from .problem import PileupProblem

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libvb2host.so")
CLI_PATH = os.path.join(PKG_DIR, "VerifyBamID")

_CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int)
_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run __graft_entry__.build()" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.vb2_host_amoeba_minimize.restype = ctypes.c_double
        lib.vb2_host_amoeba_minimize.argtypes = [_CB, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
                                                 ctypes.POINTER(ctypes.c_long)]
        lib.vb2_host_load.restype = ctypes.c_void_p
        lib.vb2_host_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        lib.vb2_host_summary.restype = ctypes.c_int
        lib.vb2_host_summary.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int64),
                                         ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]
        lib.vb2_host_copy.restype = ctypes.c_int
        lib.vb2_host_copy.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 8
        lib.vb2_host_cohort_selftest.restype = ctypes.c_long
        lib.vb2_host_cohort_selftest.argtypes = [_CB, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                 ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.vb2_host_free.restype = None
        lib.vb2_host_free.argtypes = [ctypes.c_void_p]
        lib.vb2_host_read_vcf.restype = ctypes.c_void_p
        lib.vb2_host_read_vcf.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                          ctypes.c_char_p, ctypes.c_int]
        lib.vb2_host_vcf_copy.restype = None
        lib.vb2_host_vcf_copy.argtypes = [ctypes.c_void_p] * 5
        lib.vb2_host_vcf_chrom.restype = ctypes.c_char_p
        lib.vb2_host_vcf_chrom.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.vb2_host_vcf_free.restype = None
        lib.vb2_host_vcf_free.argtypes = [ctypes.c_void_p]
        _lib = lib
    return _lib


def amoeba_minimize(fn: Callable[[np.ndarray], float], start: Sequence[float], ftol: float = 1e-8) -> Tuple[float, np.ndarray, int]:
    """AmoebaMinimizer::Reset(dim) + Minimize(ftol) on a Python callable: (fmin, point, cycleCount)."""
    lib = load_library()
    pt = np.ascontiguousarray(start, dtype=np.float64).copy()
    cb = _CB(lambda user, v, n: float(fn(np.ctypeslib.as_array(v, shape=(n,)).copy())))
    cyc = ctypes.c_long()
    r = lib.vb2_host_amoeba_minimize(cb, None, pt.size, pt.ctypes.data, float(ftol), ctypes.byref(cyc))
    return float(r), pt, int(cyc.value)


def cohort_selftest(fn: Callable[[np.ndarray], float], starts: np.ndarray, ftol: float = 1e-8):
    """n samples minimise f_i(x) = fn(x - 0.1 i) concurrently through one CohortCoordinator (host launcher).
    Returns (launches, points[n, dim], fmin[n], cycles[n])."""
    lib = load_library()
    st = np.ascontiguousarray(starts, dtype=np.float64)
    n, dim = st.shape
    pts = np.empty((n, dim)); fmin = np.empty(n); cyc = np.empty(n, dtype=np.int64)
    cb = _CB(lambda user, v, k: float(fn(np.ctypeslib.as_array(v, shape=(k,)).copy())))
    launches = lib.vb2_host_cohort_selftest(cb, None, n, dim, st.ctypes.data, float(ftol), pts.ctypes.data, fmin.ctypes.data,
                                            cyc.ctypes.data)
    return int(launches), pts, fmin, cyc


def load_problem(svd_prefix: str, pileup: str, n_pc: int = 2, disable_sanity: bool = False,
                 known_af: Optional[str] = None) -> Tuple[PileupProblem, dict]:
    """Read <prefix>.UD/.mu/.bed and a pileup with the CLI's own C++ code.  Returns (problem, summary)."""
    lib = load_library()
    h = lib.vb2_host_load(svd_prefix.encode(), pileup.encode(), int(n_pc), int(disable_sanity),
                          known_af.encode() if known_af else None)
    if not h:
        raise RuntimeError("vb2_host_load failed for %s / %s" % (svd_prefix, pileup))
    try:
        avg, sd = ctypes.c_double(), ctypes.c_double()
        nb, eff, nm = ctypes.c_long(), ctypes.c_int(), ctypes.c_uint32()
        n_info, n_reads, ok = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
        lib.vb2_host_summary(h, ctypes.byref(avg), ctypes.byref(sd), ctypes.byref(nb), ctypes.byref(eff), ctypes.byref(nm),
                             ctypes.byref(n_info), ctypes.byref(n_reads), ctypes.byref(ok))
        m, r = nm.value, n_reads.value
        idx = np.empty(m, np.int32); alt = np.empty(m, np.uint8)
        kaf = np.empty(m, np.float64) if known_af else None
        off = np.empty(n_info.value + 1, np.int64)
        bases = np.empty(r, np.uint8); quals = np.empty(r, np.uint8)
        ud = np.empty((m, n_pc), np.float64); means = np.empty(m, np.float64)
        lib.vb2_host_copy(h, idx.ctypes.data, alt.ctypes.data, kaf.ctypes.data if kaf is not None else None,
                          off.ctypes.data, bases.ctypes.data, quals.ctypes.data, ud.ctypes.data, means.ctypes.data)
        prob = PileupProblem(ud, means, idx, alt, off, bases, quals, kaf, bool(disable_sanity), avg.value, sd.value, m)
        return prob, {"avg_depth": avg.value, "sd_depth": sd.value, "num_bases": nb.value,
                      "effective_num_site": eff.value, "num_marker": m, "sanity_ok": bool(ok.value)}
    finally:
        lib.vb2_host_free(h)


def read_vcf(path: str, include_chr: Sequence[str] = ()):
    """SVDcalculator::ReadVcf of the host side (csrc/svd_panel.cpp): dict(genotype int8 [markers][samples], chrom, pos, ref, alt)."""
    lib = load_library()
    nm, ns = ctypes.c_int(), ctypes.c_int()
    err = ctypes.create_string_buffer(512)
    h = lib.vb2_host_read_vcf(path.encode(), ",".join(include_chr).encode(), ctypes.byref(nm), ctypes.byref(ns), err, 512)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        g = np.zeros((nm.value, ns.value), np.int8)
        pos = np.zeros(nm.value, np.int32)
        ref = np.zeros(nm.value, np.uint8)
        alt = np.zeros(nm.value, np.uint8)
        lib.vb2_host_vcf_copy(h, g.ctypes.data, pos.ctypes.data, ref.ctypes.data, alt.ctypes.data)
        chrom = [lib.vb2_host_vcf_chrom(h, i).decode() for i in range(nm.value)]
    finally:
        lib.vb2_host_vcf_free(h)
    return {"genotype": g, "chrom": chrom, "pos": pos, "ref": ref.tobytes().decode(), "alt": alt.tobytes().decode()}
