"""ctypes binding of libvb2llk.so (include/vb2_llk.h) -- the call a Python user makes.

`LLKEngine.compute_mix_llks(pc_contam, pc_intended, alpha)` has the signature and the value
contract of the reference's FullLLKFunc::ComputeMixLLKs (ContaminationEstimator.h:194-195):
it returns +LLK and the caller negates.

There is no CPU fallback: if the CUDA library is missing or no device is usable this module
raises -- it never routes through oracle/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from .problem import PileupProblem

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VB2_LLK_LIBRARY") or os.path.join(PKG_DIR, "libvb2llk.so")  # (override: A/B builds)

VB2_OK = 0
VB2_PANEL_FP32 = 0
VB2_PANEL_FP64 = 1
VB2_FLAG_NO_SPIN = 1
VB2_FLAG_BATCHED = 2
VB2_MAX_PC = 16
VB2_MAX_BATCH = 4096
VB2_MIN_MAX_DIM = 9

VB2_ERR_UNSUPPORTED = 6
_STATUS = {1: "VB2_ERR_INVALID", 2: "VB2_ERR_NO_DEVICE", 3: "VB2_ERR_CUDA", 4: "VB2_ERR_NOMEM", 5: "VB2_ERR_TIMEOUT",
           6: "VB2_ERR_UNSUPPORTED"}

# every symbol include/vb2_llk.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = ("vb2_abi_version", "vb2_device_count", "vb2_llk_warmup", "vb2_llk_create", "vb2_llk_destroy", "vb2_llk_get_info",
               "vb2_llk_eval", "vb2_llk_eval_begin", "vb2_llk_eval_end", "vb2_llk_eval_batch", "vb2_llk_eval_batch_device", "vb2_llk_eval_many", "vb2_llk_eval_many_device",
               "vb2_llk_sync", "vb2_llk_batch_plan", "vb2_last_error", "vb2_llk_pack_host", "vb2_llk_pack_free",
               "vb2_llk_time_device", "vb2_llk_time_device_many", "vb2_llk_time_host", "vb2_llk_trace",
               "vb2_llk_session_begin", "vb2_llk_session_end", "vb2_llk_minimize",
               "vb2_peer_create", "vb2_peer_connect", "vb2_peer_destroy", "vb2_llk_eval_many_device_peer",
               "vb2_panel_create", "vb2_panel_destroy", "vb2_ingest_parse", "vb2_ingest_flatten", "vb2_ingest_destroy",
               "vb2_llk_debug_image")


class VB2Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s: %s" % (_STATUS.get(code, "VB2_ERR_%d" % code), msg))
        self.code = code


class _Desc(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_marker", ctypes.c_uint32), ("n_pc", ctypes.c_uint32),
                ("ud_stride", ctypes.c_uint32), ("ud", ctypes.c_void_p), ("means", ctypes.c_void_p),
                ("base_info_index", ctypes.c_void_p), ("alt_base", ctypes.c_void_p), ("known_af", ctypes.c_void_p),
                ("info_offset", ctypes.c_void_p), ("bases", ctypes.c_void_p), ("quals", ctypes.c_void_p),
                ("sanity_disabled", ctypes.c_int32), ("device", ctypes.c_int32),
                ("avg_depth", ctypes.c_double), ("sd_depth", ctypes.c_double),
                ("min_af", ctypes.c_double), ("max_af", ctypes.c_double),
                ("panel_dtype", ctypes.c_int32), ("flags", ctypes.c_uint32),
                ("shard_rank", ctypes.c_uint32), ("shard_count", ctypes.c_uint32), ("stream", ctypes.c_void_p),
                ("n_info", ctypes.c_int64)]


class _Info(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_pc", ctypes.c_uint32),
                ("markers_used", ctypes.c_uint64), ("reads_used", ctypes.c_uint64),
                ("reads_streamed", ctypes.c_uint64), ("reads_folded", ctypes.c_uint64),
                ("algorithmic_bytes", ctypes.c_uint64), ("device_bytes", ctypes.c_uint64),
                ("n_slices", ctypes.c_uint32), ("grid_x", ctypes.c_uint32), ("block_threads", ctypes.c_uint32),
                ("smem_bytes", ctypes.c_uint32), ("device", ctypes.c_int32), ("sm_count", ctypes.c_int32),
                ("log_other_const", ctypes.c_double)]


class _PackedView(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_pc", ctypes.c_uint32), ("n_used", ctypes.c_uint32),
                ("n_slices", ctypes.c_uint32), ("grid_x", ctypes.c_uint32), ("n_bins", ctypes.c_uint32),
                ("conc_rounds", ctypes.c_uint32), ("n_rounds", ctypes.c_uint32), ("max_stride", ctypes.c_uint32),
                ("known_af", ctypes.c_uint32), ("panel_elem", ctypes.c_uint32),
                ("off_ud", ctypes.c_uint32), ("off_mu", ctypes.c_uint32), ("off_kaf", ctypes.c_uint32),
                ("off_diag", ctypes.c_uint32), ("off_words", ctypes.c_uint32),
                ("reads_used", ctypes.c_uint64), ("reads_streamed", ctypes.c_uint64),
                ("reads_folded", ctypes.c_uint64), ("blob_bytes", ctypes.c_uint64),
                ("log_other_const", ctypes.c_double),
                ("blob", ctypes.POINTER(ctypes.c_uint8)), ("rounds", ctypes.POINTER(ctypes.c_uint32)),
                ("marker_index", ctypes.POINTER(ctypes.c_uint32)), ("owner", ctypes.c_void_p)]


class _Model(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("dim", ctypes.c_uint32),
                ("pc1_from", ctypes.c_int32 * VB2_MAX_PC), ("pc2_from", ctypes.c_int32 * VB2_MAX_PC),
                ("alpha_from", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("pc1_fixed", ctypes.c_double * VB2_MAX_PC), ("pc2_fixed", ctypes.c_double * VB2_MAX_PC),
                ("alpha_fixed", ctypes.c_double)]


class _MinResult(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("converged", ctypes.c_int32), ("fmin", ctypes.c_double),
                ("point", ctypes.c_double * VB2_MIN_MAX_DIM), ("evals", ctypes.c_int64), ("cycle_count", ctypes.c_int64),
                ("llk1", ctypes.c_double), ("improved", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("best_pc_contam", ctypes.c_double * 4), ("best_pc_intended", ctypes.c_double * 4),
                ("best_alpha", ctypes.c_double)]


class _PanelDesc(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_marker", ctypes.c_uint32), ("n_pc", ctypes.c_uint32),
                ("ud_stride", ctypes.c_uint32), ("ud", ctypes.c_void_p), ("means", ctypes.c_void_p),
                ("chrom_id", ctypes.c_void_p), ("pos", ctypes.c_void_p), ("alt_base", ctypes.c_void_p),
                ("chrom_names", ctypes.c_char_p), ("n_chrom", ctypes.c_uint32), ("device", ctypes.c_int32)]


class _IngestInfo(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_lines", ctypes.c_uint32), ("n_matched", ctypes.c_uint32),
                ("pad_", ctypes.c_uint32), ("num_bases", ctypes.c_uint64), ("row_depth", ctypes.POINTER(ctypes.c_int32))]


class _FlattenDesc(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("stream", ctypes.c_void_p),
                ("flags", ctypes.c_uint32), ("panel_dtype", ctypes.c_int32), ("min_af", ctypes.c_double),
                ("max_af", ctypes.c_double), ("sanity_disabled", ctypes.c_int32), ("shard_rank", ctypes.c_uint32),
                ("shard_count", ctypes.c_uint32), ("pad_", ctypes.c_uint32), ("avg_depth", ctypes.c_double),
                ("sd_depth", ctypes.c_double)]


_lib = None


def build_library(quiet: bool = True) -> None:
    """Compile libvb2llk.so (and the CLI) in-tree with nvcc for sm_100a."""
    subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc"), "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def load_library() -> ctypes.CDLL:
    """Load the CUDA engine.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the likelihood path has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.vb2_abi_version.restype = ctypes.c_int
    lib.vb2_device_count.restype = ctypes.c_int
    lib.vb2_last_error.restype = ctypes.c_char_p
    lib.vb2_last_error.argtypes = [ctypes.c_void_p]
    lib.vb2_llk_create.restype = ctypes.c_int
    lib.vb2_llk_create.argtypes = [ctypes.POINTER(_Desc), ctypes.POINTER(ctypes.c_void_p)]
    lib.vb2_llk_destroy.restype = None
    lib.vb2_llk_destroy.argtypes = [ctypes.c_void_p]
    lib.vb2_llk_get_info.restype = ctypes.c_int
    lib.vb2_llk_get_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Info)]
    lib.vb2_llk_eval.restype = ctypes.c_int
    lib.vb2_llk_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                 ctypes.POINTER(ctypes.c_double)]
    lib.vb2_llk_eval_begin.restype = ctypes.c_int
    lib.vb2_llk_eval_begin.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
    lib.vb2_llk_eval_end.restype = ctypes.c_int
    lib.vb2_llk_eval_end.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    for name in ("vb2_llk_eval_batch", "vb2_llk_eval_batch_device"):
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.vb2_llk_eval_many.restype = ctypes.c_int
    lib.vb2_llk_eval_many.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p]
    lib.vb2_llk_eval_many_device.restype = ctypes.c_int
    lib.vb2_llk_eval_many_device.argtypes = lib.vb2_llk_eval_many.argtypes
    if hasattr(lib, "vb2_llk_trace"):  # (absent from older A/B builds loaded through VB2_LLK_LIBRARY)
        lib.vb2_llk_trace.restype = ctypes.c_int
        lib.vb2_llk_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p,
                                      ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_double)]
    for name in ("vb2_llk_session_begin", "vb2_llk_session_end"):
        if hasattr(lib, name):
            getattr(lib, name).restype = ctypes.c_int
            getattr(lib, name).argtypes = [ctypes.c_void_p]
    lib.vb2_panel_create.restype = ctypes.c_int
    lib.vb2_panel_create.argtypes = [ctypes.POINTER(_PanelDesc), ctypes.POINTER(ctypes.c_void_p)]
    lib.vb2_panel_destroy.restype = None
    lib.vb2_panel_destroy.argtypes = [ctypes.c_void_p]
    lib.vb2_ingest_parse.restype = ctypes.c_int
    lib.vb2_ingest_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(_IngestInfo)]
    lib.vb2_ingest_flatten.restype = ctypes.c_int
    lib.vb2_ingest_flatten.argtypes = [ctypes.c_void_p, ctypes.POINTER(_FlattenDesc), ctypes.POINTER(ctypes.c_void_p)]
    lib.vb2_ingest_destroy.restype = None
    lib.vb2_ingest_destroy.argtypes = [ctypes.c_void_p]
    lib.vb2_llk_debug_image.restype = ctypes.c_int
    lib.vb2_llk_debug_image.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    lib.vb2_peer_create.restype = ctypes.c_int
    lib.vb2_peer_create.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
    lib.vb2_peer_connect.restype = ctypes.c_int
    lib.vb2_peer_connect.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.vb2_peer_destroy.restype = None
    lib.vb2_peer_destroy.argtypes = [ctypes.c_void_p]
    lib.vb2_llk_eval_many_device_peer.restype = ctypes.c_int
    lib.vb2_llk_eval_many_device_peer.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.vb2_llk_minimize.restype = ctypes.c_int
    lib.vb2_llk_minimize.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Model), ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int64, ctypes.c_double, ctypes.POINTER(_MinResult)]
    lib.vb2_llk_sync.restype = ctypes.c_int
    lib.vb2_llk_sync.argtypes = [ctypes.c_void_p]
    lib.vb2_llk_batch_plan.restype = ctypes.c_int
    lib.vb2_llk_batch_plan.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int)]
    lib.vb2_llk_time_device.restype = ctypes.c_int
    lib.vb2_llk_time_device.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                        ctypes.POINTER(ctypes.c_float)]
    lib.vb2_llk_time_device_many.restype = ctypes.c_int
    lib.vb2_llk_time_device_many.argtypes = lib.vb2_llk_time_device.argtypes
    lib.vb2_llk_time_host.restype = ctypes.c_int
    lib.vb2_llk_time_host.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.vb2_llk_pack_host.restype = ctypes.c_int
    lib.vb2_llk_pack_host.argtypes = [ctypes.POINTER(_Desc), ctypes.c_uint32, ctypes.POINTER(_PackedView)]
    lib.vb2_llk_pack_free.restype = None
    lib.vb2_llk_pack_free.argtypes = [ctypes.POINTER(_PackedView)]
    if lib.vb2_abi_version() != 2:
        raise RuntimeError("libvb2llk.so ABI version mismatch")
    _lib = lib
    return lib


def device_count() -> int:
    return int(load_library().vb2_device_count())


def _f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and out.shape != shape:
        out = out.reshape(shape)
    return out


def make_desc(problem: PileupProblem, device: int = 0, panel_dtype: int = VB2_PANEL_FP32, shard_rank: int = 0,
              shard_count: int = 1, stream: Optional[int] = None, spin: bool = True, min_af: float = 0.0,
              max_af: float = 0.0, batched: bool = False) -> _Desc:
    """vb2_llk_desc over the numpy arrays of `problem` (which must stay alive during the call)."""
    d = _Desc()
    d.struct_size = ctypes.sizeof(_Desc)
    d.n_marker = problem.n_marker
    d.n_pc = problem.n_pc
    d.ud_stride = problem.n_pc
    d.ud = problem.ud.ctypes.data
    d.means = problem.means.ctypes.data
    d.base_info_index = problem.base_info_index.ctypes.data
    d.alt_base = problem.alt_base.ctypes.data
    d.known_af = problem.known_af.ctypes.data if problem.known_af is not None else None
    d.info_offset = problem.info_offset.ctypes.data
    d.bases = problem.bases.ctypes.data if problem.bases.size else None
    d.quals = problem.quals.ctypes.data if problem.quals.size else None
    d.sanity_disabled = int(problem.sanity_disabled)
    d.device = int(device)
    d.avg_depth = float(problem.avg_depth)
    d.sd_depth = float(problem.sd_depth)
    d.min_af, d.max_af = float(min_af), float(max_af)
    d.panel_dtype = int(panel_dtype)
    d.flags = (0 if spin else VB2_FLAG_NO_SPIN) | (VB2_FLAG_BATCHED if batched else 0)
    d.shard_rank, d.shard_count = int(shard_rank), int(shard_count)
    d.stream = stream
    d.n_info = max(0, int(problem.info_offset.size) - 1)
    return d


def pack_host(problem: PileupProblem, shard_rank: int = 0, shard_count: int = 1, max_ctas: int = 148,
              panel_dtype: int = VB2_PANEL_FP64, batched: bool = False) -> dict:
    """Host-only: the flattened image vb2_llk_create would upload, as numpy copies (no CUDA call).

    Returns the scalars of vb2_packed_view plus `blob` (uint8), `rounds` (list of dicts with base, stride,
    first_bin, count, rows) and `marker_index`."""
    lib = load_library()
    d = make_desc(problem, shard_rank=shard_rank, shard_count=shard_count, panel_dtype=panel_dtype, batched=batched)
    v = _PackedView()
    v.struct_size = ctypes.sizeof(_PackedView)
    rc = lib.vb2_llk_pack_host(ctypes.byref(d), int(max_ctas), ctypes.byref(v))
    if rc != VB2_OK:
        raise VB2Error(rc, "vb2_llk_pack_host failed")
    try:
        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        out = {name: getattr(v, name) for name, _ in _PackedView._fields_
               if name not in ("struct_size", "blob", "rounds", "marker_index", "owner")}
        out["blob"] = arr(v.blob, v.blob_bytes, np.uint8)
        rr = arr(v.rounds, 6 * v.n_rounds, np.uint32).reshape(-1, 6)
        out["rounds"] = [{"base": int(r[0]) | (int(r[1]) << 32), "stride": int(r[2]), "first_bin": int(r[3]),
                          "count": int(r[4]), "rows": int(r[5])} for r in rr]
        out["marker_index"] = arr(v.marker_index, 32 * v.n_slices, np.uint32)
        return out
    finally:
        lib.vb2_llk_pack_free(ctypes.byref(v))


class DevicePanel:
    """A reference panel resident on the device for the device-side pileup ingest (vb2_panel_create)."""

    def __init__(self, ud, means, chrom: Sequence[str], pos, alt_base, device: int = 0):
        self._lib = load_library()
        self.ud = np.ascontiguousarray(ud, dtype=np.float64)
        self.means = np.ascontiguousarray(means, dtype=np.float64)
        names = sorted(set(chrom))
        index = {c: i for i, c in enumerate(names)}
        self.chrom_id = np.array([index[c] for c in chrom], dtype=np.uint16)
        self.pos = np.ascontiguousarray(pos, dtype=np.int32)
        self.alt = np.ascontiguousarray(alt_base, dtype=np.uint8)
        self.n_marker, self.n_pc = self.ud.shape
        self.device = int(device)
        d = _PanelDesc()
        d.struct_size = ctypes.sizeof(_PanelDesc)
        d.n_marker, d.n_pc, d.ud_stride = self.n_marker, self.n_pc, self.n_pc
        d.ud, d.means = self.ud.ctypes.data, self.means.ctypes.data
        d.chrom_id, d.pos, d.alt_base = self.chrom_id.ctypes.data, self.pos.ctypes.data, self.alt.ctypes.data
        self._names = b"".join(c.encode() + b"\0" for c in names)
        d.chrom_names = self._names
        d.n_chrom = len(names)
        d.device = self.device
        self._panel = ctypes.c_void_p()
        rc = self._lib.vb2_panel_create(ctypes.byref(d), ctypes.byref(self._panel))
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(None).decode())

    def ingest(self, text: bytes, sanity_disabled: bool = False, panel_dtype: int = VB2_PANEL_FP32, batched: bool = False,
               stats=None) -> "LLKEngine":
        """Pileup text -> a resident engine, parsed and flattened on the device.  `stats(info) -> (avg_depth, sd_depth)`
        is the caller's marker sanity check (default: the reference's IsSanityCheckOK arithmetic); raises VB2Error with
        code VB2_ERR_UNSUPPORTED for text that needs the host reader."""
        ing = ctypes.c_void_p()
        info = _IngestInfo()
        info.struct_size = ctypes.sizeof(_IngestInfo)
        rc = self._lib.vb2_ingest_parse(self._panel, text, len(text), ctypes.byref(ing), ctypes.byref(info))
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(None).decode())
        try:
            depth = np.ctypeslib.as_array(info.row_depth, shape=(self.n_marker,)).copy() if self.n_marker else np.zeros(0, np.int32)
            summary = {"n_lines": info.n_lines, "n_matched": info.n_matched, "num_bases": info.num_bases, "row_depth": depth}
            if stats is not None:
                avg, sd = stats(summary)
            else:   # ContaminationEstimator.cpp:543-565 (int multiply, effectiveNumSite = matched lines)
                avg = info.num_bases / info.n_matched if info.n_matched else float("nan")
                have = depth[depth >= 0].astype(np.int64)
                sq = float(np.sum((have.astype(np.int32) * have.astype(np.int32)).astype(np.int64)))
                sd = float(np.sqrt(sq / info.n_matched - avg * avg)) if info.n_matched else float("nan")
            f = _FlattenDesc()
            f.struct_size = ctypes.sizeof(_FlattenDesc)
            f.device = self.device
            f.stream = None
            f.flags = VB2_FLAG_BATCHED if batched else 0
            f.panel_dtype = int(panel_dtype)
            f.sanity_disabled = int(sanity_disabled)
            f.shard_rank, f.shard_count = 0, 1
            f.avg_depth, f.sd_depth = float(avg), float(sd)
            ctx = ctypes.c_void_p()
            rc = self._lib.vb2_ingest_flatten(ing, ctypes.byref(f), ctypes.byref(ctx))
            if rc != VB2_OK:
                raise VB2Error(rc, self._lib.vb2_last_error(None).decode())
        finally:
            self._lib.vb2_ingest_destroy(ing)
        eng = LLKEngine.__new__(LLKEngine)
        eng._lib, eng._ctx, eng.problem, eng.n_pc = self._lib, ctx, None, self.n_pc
        eng.ingest_summary = dict(summary, avg_depth=float(avg), sd_depth=float(sd))
        return eng

    def close(self) -> None:
        if self._panel:
            self._lib.vb2_panel_destroy(self._panel)
            self._panel = ctypes.c_void_p()


class LLKEngine:
    """One sample resident in the HBM of one GPU (or one marker shard of it)."""

    def __init__(self, problem: PileupProblem, device: int = 0, panel_dtype: int = VB2_PANEL_FP32,
                 shard_rank: int = 0, shard_count: int = 1, stream: Optional[int] = None, spin: bool = True,
                 min_af: float = 0.0, max_af: float = 0.0, batched: bool = False):
        self._lib = load_library()
        self._ctx = ctypes.c_void_p()
        self.problem = problem
        d = make_desc(problem, device, panel_dtype, shard_rank, shard_count, stream, spin, min_af, max_af, batched)
        rc = self._lib.vb2_llk_create(ctypes.byref(d), ctypes.byref(self._ctx))
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(None).decode())
        self.n_pc = problem.n_pc

    # -- lifetime --------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.vb2_llk_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(self._ctx).decode())

    # -- the reference interface -------------------------------------------------------------
    def compute_mix_llks(self, pc_contam: Sequence[float], pc_intended: Sequence[float], alpha: float) -> float:
        """ComputeMixLLKs(tPC1, tPC2, alpha): tPC1 = contaminating, tPC2 = intended sample PCs."""
        a, b = _f64(pc_contam), _f64(pc_intended)
        if a.size != self.n_pc or b.size != self.n_pc:
            raise ValueError("PC vectors must have n_pc=%d entries" % self.n_pc)
        out = ctypes.c_double()
        self._check(self._lib.vb2_llk_eval(self._ctx, a.ctypes.data, b.ctypes.data, float(alpha), ctypes.byref(out)))
        return float(out.value)

    def eval_begin(self, pc_contam: Sequence[float], pc_intended: Sequence[float], alpha: float) -> None:
        """Launch one evaluation without waiting (see eval_end): lets one thread drive several shards."""
        a, b = _f64(pc_contam), _f64(pc_intended)
        if a.size != self.n_pc or b.size != self.n_pc:
            raise ValueError("PC vectors must have n_pc=%d entries" % self.n_pc)
        self._check(self._lib.vb2_llk_eval_begin(self._ctx, a.ctypes.data, b.ctypes.data, float(alpha)))

    def eval_end(self) -> float:
        out = ctypes.c_double()
        self._check(self._lib.vb2_llk_eval_end(self._ctx, ctypes.byref(out)))
        return float(out.value)

    def eval_batch(self, pc_contam, pc_intended, alphas) -> np.ndarray:
        al = _f64(alphas).ravel()
        n = al.size
        a, b = _f64(pc_contam, (n, self.n_pc)), _f64(pc_intended, (n, self.n_pc))
        out = np.empty(n, dtype=np.float64)
        self._check(self._lib.vb2_llk_eval_batch(self._ctx, n, a.ctypes.data, b.ctypes.data, al.ctypes.data,
                                                 out.ctypes.data))
        return out

    def eval_batch_device(self, pc_contam, pc_intended, alphas, d_out_ptr: int) -> None:
        """Asynchronous; results stay in device memory at `d_out_ptr` (n doubles)."""
        al = _f64(alphas).ravel()
        n = al.size
        a, b = _f64(pc_contam, (n, self.n_pc)), _f64(pc_intended, (n, self.n_pc))
        self._check(self._lib.vb2_llk_eval_batch_device(self._ctx, n, a.ctypes.data, b.ctypes.data, al.ctypes.data,
                                                        ctypes.c_void_p(d_out_ptr)))

    def sync(self) -> None:
        self._check(self._lib.vb2_llk_sync(self._ctx))

    def batch_plan(self, n: int) -> dict:
        """How a batched call over n evaluations of samples shaped like this one is launched (nothing is launched)."""
        flow, launches, jobs = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self._lib.vb2_llk_batch_plan(self._ctx, int(n), ctypes.byref(flow), ctypes.byref(launches), ctypes.byref(jobs)))
        return {"kernel": "llk_flow_kernel" if flow.value else "llk_stream_kernel", "kernel_launches": launches.value,
                "jobs_per_launch": jobs.value}

    def session_begin(self) -> None:
        """Resident kernel with the sample in shared memory: compute_mix_llks rings a doorbell until session_end()."""
        self._check(self._lib.vb2_llk_session_begin(self._ctx))

    def session_end(self) -> None:
        self._check(self._lib.vb2_llk_session_end(self._ctx))

    def minimize(self, start, pc1_from, pc2_from, alpha_from: int, pc1_fixed=None, pc2_fixed=None, alpha_fixed: float = 0.0,
                 scale: float = 1.0, ftol: float = 1e-8, cycle_max: int = 50000, llk1: float = 1e300) -> dict:
        """AmoebaMinimizer::Minimize on the device (include/vb2_llk.h, vb2_llk_minimize); needs session_begin().
        pc1_from / pc2_from: per PC the index into the simplex vector, or -1 = the fixed value; alpha_from likewise
        (alpha = InvLogit(v[alpha_from])).  Returns the fields of vb2_llk_min_result."""
        v = _f64(start).ravel()
        m = _Model()
        m.struct_size = ctypes.sizeof(_Model)
        m.dim = v.size
        for k in range(VB2_MAX_PC):
            m.pc1_from[k] = int(pc1_from[k]) if k < len(pc1_from) else -1
            m.pc2_from[k] = int(pc2_from[k]) if k < len(pc2_from) else -1
            m.pc1_fixed[k] = float(pc1_fixed[k]) if pc1_fixed is not None and k < len(pc1_fixed) else 0.0
            m.pc2_fixed[k] = float(pc2_fixed[k]) if pc2_fixed is not None and k < len(pc2_fixed) else 0.0
        m.alpha_from = int(alpha_from)
        m.alpha_fixed = float(alpha_fixed)
        r = _MinResult()
        r.struct_size = ctypes.sizeof(_MinResult)
        self._check(self._lib.vb2_llk_minimize(self._ctx, ctypes.byref(m), v.ctypes.data, float(scale), float(ftol),
                                               int(cycle_max), float(llk1), ctypes.byref(r)))
        return {"converged": bool(r.converged), "fmin": r.fmin, "point": [r.point[i] for i in range(v.size)], "evals": int(r.evals),
                "cycle_count": int(r.cycle_count), "llk1": r.llk1, "improved": bool(r.improved),
                "best_pc_contam": [r.best_pc_contam[i] for i in range(self.n_pc)],
                "best_pc_intended": [r.best_pc_intended[i] for i in range(self.n_pc)], "best_alpha": r.best_alpha}

    def trace(self, pc_contam, pc_intended, alpha: float):
        """One evaluation with the kernel's stage clock on: (llk, stamps[cta][VB2_TRACE_SLOTS]) (include/vb2_llk.h)."""
        a = np.ascontiguousarray(pc_contam, dtype=np.float64)
        b = np.ascontiguousarray(pc_intended, dtype=np.float64)
        stamps = np.zeros((1024, 16), dtype=np.uint64)
        n = ctypes.c_uint32()
        llk = ctypes.c_double()
        self._check(self._lib.vb2_llk_trace(self._ctx, a.ctypes.data, b.ctypes.data, float(alpha), stamps.ctypes.data,
                                            1024, ctypes.byref(n), ctypes.byref(llk)))
        return llk.value, stamps[:n.value]

    def debug_image(self, n_bytes: int) -> np.ndarray:
        """The first n_bytes of the image as it sits in device memory (tests)."""
        out = np.empty(int(n_bytes), dtype=np.uint8)
        self._check(self._lib.vb2_llk_debug_image(self._ctx, out.ctypes.data, int(n_bytes)))
        return out

    def info(self) -> dict:
        i = _Info()
        i.struct_size = ctypes.sizeof(_Info)
        self._check(self._lib.vb2_llk_get_info(self._ctx, ctypes.byref(i)))
        return {name: getattr(i, name) for name, _ in _Info._fields_ if name != "struct_size"}


def eval_many(engines: List[LLKEngine], pc_contam, pc_intended, alphas) -> np.ndarray:
    """One evaluation of each of n different samples (same device, same n_pc) in one launch."""
    n = len(engines)
    k = engines[0].n_pc
    al = _f64(alphas).ravel()
    a, b = _f64(pc_contam, (n, k)), _f64(pc_intended, (n, k))
    arr = (ctypes.c_void_p * n)(*[e._ctx for e in engines])
    out = np.empty(n, dtype=np.float64)
    lib = load_library()
    rc = lib.vb2_llk_eval_many(arr, n, a.ctypes.data, b.ctypes.data, al.ctypes.data, out.ctypes.data)
    if rc != VB2_OK:
        raise VB2Error(rc, lib.vb2_last_error(engines[0]._ctx).decode())
    return out


def context_array(engines: List[LLKEngine]):
    """The ctypes array of context handles the *_many calls take (build it once for a list that is used again)."""
    return (ctypes.c_void_p * len(engines))(*[e._ctx for e in engines])


def eval_many_device(engines: List[LLKEngine], pc_contam, pc_intended, alphas, d_out_ptr: int, ctx_array=None) -> None:
    """Asynchronous eval_many: results stay in device memory at `d_out_ptr` (len(engines) doubles).  `ctx_array`:
    context_array(engines) built beforehand (only its first len(engines) entries are used)."""
    n = len(engines)
    k = engines[0].n_pc
    al = _f64(alphas).ravel()
    a, b = _f64(pc_contam, (n, k)), _f64(pc_intended, (n, k))
    arr = ctx_array if ctx_array is not None else context_array(engines)
    lib = load_library()
    rc = lib.vb2_llk_eval_many_device(arr, n, a.ctypes.data, b.ctypes.data, al.ctypes.data, ctypes.c_void_p(d_out_ptr))
    if rc != VB2_OK:
        raise VB2Error(rc, lib.vb2_last_error(engines[0]._ctx).decode())


class PeerReduce:
    """The cross-GPU sum fused into the kernels (include/vb2_llk.h, vb2_peer_*): one process per GPU, buffers exchanged as
    CUDA IPC handles through `exchange(handle_bytes) -> list of every rank's handle bytes, in rank order`."""

    def __init__(self, device: int, rank: int, world: int, exchange):
        self._lib = load_library()
        self._peer = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        rc = self._lib.vb2_peer_create(int(device), int(rank), int(world), ctypes.byref(self._peer), handle)
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(None).decode())
        handles = exchange(bytes(handle))
        if len(handles) != world or any(len(h) != 64 for h in handles):
            raise ValueError("exchange() must return one 64-byte handle per rank")
        blob = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        rc = self._lib.vb2_peer_connect(self._peer, blob)
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(None).decode())

    def eval_many_device(self, engines: List[LLKEngine], pc_contam, pc_intended, alphas, d_out_ptr: int, ctx_array=None) -> None:
        """Every rank's shard of the same evaluations; the sums over all shards land in device memory at `d_out_ptr`."""
        n = len(engines)
        k = engines[0].n_pc
        al = _f64(alphas).ravel()
        a, b = _f64(pc_contam, (n, k)), _f64(pc_intended, (n, k))
        arr = ctx_array if ctx_array is not None else context_array(engines)
        rc = self._lib.vb2_llk_eval_many_device_peer(arr, n, a.ctypes.data, b.ctypes.data, al.ctypes.data, self._peer,
                                                     ctypes.c_void_p(d_out_ptr))
        if rc != VB2_OK:
            raise VB2Error(rc, self._lib.vb2_last_error(engines[0]._ctx).decode())

    def close(self) -> None:
        if self._peer:
            self._lib.vb2_peer_destroy(self._peer)
            self._peer = ctypes.c_void_p()


def time_device(engines: List[LLKEngine], warmup: int, steps: int, pc_contam, pc_intended, alpha: float) -> float:
    """Milliseconds (CUDA events on the engines' stream) for `steps` back-to-back evaluations issued from C,
    step i on engines[i % n]."""
    lib = load_library()
    n = len(engines)
    arr = (ctypes.c_void_p * n)(*[e._ctx for e in engines])
    a, b = _f64(pc_contam), _f64(pc_intended)
    ms = ctypes.c_float()
    rc = lib.vb2_llk_time_device(arr, n, warmup, steps, a.ctypes.data, b.ctypes.data, float(alpha), ctypes.byref(ms))
    if rc != VB2_OK:
        raise VB2Error(rc, lib.vb2_last_error(engines[0]._ctx).decode())
    return float(ms.value)


def time_device_many(engines: List[LLKEngine], warmup_launches: int, launches: int, pc_contam, pc_intended,
                     alpha: float) -> float:
    """Milliseconds for `launches` launches that each evaluate every engine's sample once (len(engines) steps
    per launch, vb2_llk_eval_many's kernel), issued back to back from C."""
    lib = load_library()
    n = len(engines)
    arr = (ctypes.c_void_p * n)(*[e._ctx for e in engines])
    a, b = _f64(pc_contam), _f64(pc_intended)
    ms = ctypes.c_float()
    rc = lib.vb2_llk_time_device_many(arr, n, warmup_launches, launches, a.ctypes.data, b.ctypes.data, float(alpha),
                                      ctypes.byref(ms))
    if rc != VB2_OK:
        raise VB2Error(rc, lib.vb2_last_error(engines[0]._ctx).decode())
    return float(ms.value)


def time_host(engines: List[LLKEngine], warmup: int, steps: int, pc_contam, pc_intended, alpha: float):
    """(seconds, last LLK) for `steps` synchronous vb2_llk_eval calls with host buffers, issued from C."""
    lib = load_library()
    n = len(engines)
    arr = (ctypes.c_void_p * n)(*[e._ctx for e in engines])
    a, b = _f64(pc_contam), _f64(pc_intended)
    s, last = ctypes.c_double(), ctypes.c_double()
    rc = lib.vb2_llk_time_host(arr, n, warmup, steps, a.ctypes.data, b.ctypes.data, float(alpha), ctypes.byref(s),
                               ctypes.byref(last))
    if rc != VB2_OK:
        raise VB2Error(rc, lib.vb2_last_error(engines[0]._ctx).decode())
    return float(s.value), float(last.value)
