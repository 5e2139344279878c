"""The C-ABI library loads, exports every symbol include/vb2_llk.h declares, and fails loudly
(never falls back to a CPU path) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

import verifybamid_b200 as vb
from verifybamid_b200 import engine
from helpers import ROOT, RESULT_PILEUP, golden_problem, to_product


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vb2_llk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = vb.load_library()
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libvb2llk.so does not export %s" % name
    assert sorted(vb.ABI_SYMBOLS) == declared


def test_abi_version_and_struct_sizes():
    lib = vb.load_library()
    assert lib.vb2_abi_version() == 2
    assert ctypes.sizeof(engine._Desc) == 152          # (version 2: + n_info)
    assert ctypes.sizeof(engine._Model) == 408 and ctypes.sizeof(engine._MinResult) == 192
    assert ctypes.sizeof(engine._PanelDesc) == 72 and ctypes.sizeof(engine._FlattenDesc) == 72 and ctypes.sizeof(engine._IngestInfo) == 32
    # a descriptor with the wrong struct_size must be refused before anything is touched
    d = engine.make_desc(to_product(golden_problem(RESULT_PILEUP)))
    d.struct_size = 8
    ctx = ctypes.c_void_p()
    assert lib.vb2_llk_create(ctypes.byref(d), ctypes.byref(ctx)) == 1
    assert b"struct_size" in lib.vb2_last_error(None)


def test_library_does_not_link_the_oracle():
    # product path must not route through oracle/: no vb2o_* symbol in the shipped library
    out = os.popen("nm -D --defined-only %s" % engine.LIB_PATH).read()
    assert "vb2o_" not in out and "vb2_llk_eval" in out


@pytest.mark.skipif(vb.device_count() > 0, reason="a CUDA device is present")
def test_no_device_fails_loudly():
    with pytest.raises(vb.VB2Error) as ei:
        vb.LLKEngine(to_product(golden_problem(RESULT_PILEUP)))
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
