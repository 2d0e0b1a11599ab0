"""CPU tests of the drop-in boundary: libsqgpu.so loads, exports every symbol
that include/sqgpu.h declares, and the Python layer fails loudly (no CPU
fallback) when there is no device.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from tests.helpers import ROOT
from sequali_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sqgpu.h")).read()
    return sorted(set(re.findall(r"SQ_API\s+[\w\s\*]+?\b(sq_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) > 50
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_sizes_match_the_header():
    assert ctypes.sizeof(_lib.Meta) == 40
    assert ctypes.sizeof(_lib.NanoInfo) == 40
    assert ctypes.sizeof(_lib.ParseInfo) == 48


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.sq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import sequali_b200
    for make in (sequali_b200.QCMetrics, sequali_b200.PerTileQuality, sequali_b200.NanoStats,
                 lambda: sequali_b200.AdapterCounter(["ACGT"]),
                 sequali_b200.OverrepresentedSequences, sequali_b200.DedupEstimator,
                 sequali_b200.InsertSizeMetrics):
        with pytest.raises(_lib.SqGpuError):
            make()
    h = ctypes.c_void_p()
    assert lib.sq_ctx_create(0, ctypes.byref(h)) == _lib.SQ_E_NODEVICE
    assert b"no CPU fallback" in lib.sq_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sequali_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("tests/", ""), os.path.join(dirpath, f)
