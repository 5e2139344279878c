"""ctypes binding of libvb2svd.so (include/vb2_svd.h): the panel-construction step of `--RefVCF` on the device.
Used by the tests and tools; the product caller is the C++ CLI (csrc/svd_panel.cpp)."""
import ctypes
import os

import numpy as np

LIB_PATH = os.environ.get("VB2_SVD_LIBRARY", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvb2svd.so"))
VB2_SVD_MAX_PC = 64
_lib = None


class _Timing(ctypes.Structure):
    _fields_ = [("center_ms", ctypes.c_float), ("gram_ms", ctypes.c_float), ("eigen_ms", ctypes.c_float),
                ("ud_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


class _Desc(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("n_marker", ctypes.c_uint32),
                ("n_sample", ctypes.c_uint32), ("n_pc", ctypes.c_uint32), ("pad_", ctypes.c_uint32),
                ("genotype", ctypes.c_void_p), ("centered", ctypes.c_void_p), ("mu", ctypes.c_void_p),
                ("ud", ctypes.c_void_p), ("pc", ctypes.c_void_p), ("singular", ctypes.c_void_p),
                ("timing", ctypes.POINTER(_Timing))]


class SVDError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vb2_svd_gram failed (%d): %s" % (code, msg))
        self.code = code


def load_library():
    """The library or an exception -- never a CPU substitute."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libvb2svd.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
        lib = ctypes.CDLL(LIB_PATH)
        lib.vb2_svd_gram.restype = ctypes.c_int
        lib.vb2_svd_gram.argtypes = [ctypes.POINTER(_Desc)]
        lib.vb2_svd_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def svd_gram(matrix, n_pc, device=0):
    """matrix: int8 genotypes [M][N] (centred on the device, mu returned) or float32 already centred [M][N].
    Returns dict(mu, ud [M][n_pc], pc [N][n_pc], singular [N], timing)."""
    lib = load_library()
    a = np.ascontiguousarray(matrix)
    m, n = a.shape
    geno = a.dtype == np.int8
    if not geno:
        a = np.ascontiguousarray(a, dtype=np.float32)
    mu = np.zeros(m, np.float32)
    ud = np.zeros((m, n_pc), np.float32)
    pc = np.zeros((n, n_pc), np.float32)
    sv = np.zeros(n, np.float32)
    t = _Timing()
    d = _Desc(ctypes.sizeof(_Desc), device, m, n, n_pc, 0, a.ctypes.data if geno else None, None if geno else a.ctypes.data,
              mu.ctypes.data, ud.ctypes.data, pc.ctypes.data, sv.ctypes.data, ctypes.pointer(t))
    rc = lib.vb2_svd_gram(ctypes.byref(d))
    if rc != 0:
        raise SVDError(rc, lib.vb2_svd_last_error().decode())
    return {"mu": mu if geno else None, "ud": ud, "pc": pc, "singular": sv,
            "timing": {k: getattr(t, k) for k, _ in _Timing._fields_}}
