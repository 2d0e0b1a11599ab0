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


def test_extension_module_exports_the_reference_surface():
    """sequali_b200/ext/_qc.so (built by __graft_entry__.build()) loads without a GPU and exposes the
    names of the reference's stub file (src/sequali/_qc.pyi:21-189)."""
    import importlib
    try:
        ext = importlib.import_module("sequali_b200.ext._qc")
    except ImportError as e:
        pytest.fail(f"the CPython extension is not built: {e}")
    names = ["FastqRecordView", "FastqRecordArrayView", "FastqParser", "BamParser", "QCMetrics", "AdapterCounter",
             "PerTileQuality", "OverrepresentedSequences", "DedupEstimator", "NanoStats", "NanoStatsIterator",
             "NanoporeReadInfo", "InsertSizeMetrics", "A", "C", "G", "T", "N", "NUMBER_OF_NUCS", "NUMBER_OF_PHREDS",
             "TABLE_SIZE", "PHRED_MAX", "MAX_SEQUENCE_SIZE", "DEFAULT_END_ANCHOR_LENGTH",
             "DEFAULT_MAX_UNIQUE_FRAGMENTS", "DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS", "DEFAULT_FRAGMENT_LENGTH",
             "DEFAULT_UNIQUE_SAMPLE_EVERY", "DEFAULT_BASES_FROM_START", "DEFAULT_BASES_FROM_END",
             "DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH", "DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH",
             "DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET", "DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET",
             "INSERT_SIZE_MAX_ADAPTER_STORE_SIZE"]
    missing = [n for n in names if not hasattr(ext, n)]
    assert not missing, missing
    import sequali_b200._qc as mirror
    for n in names:
        if isinstance(getattr(ext, n), int):
            assert getattr(ext, n) == getattr(mirror, n), n
    v = ext.FastqRecordView("r1 extra", "ACGTN", "IIII#", b"RGZa\0")
    assert (v.name(), v.sequence(), v.qualities(), v.tags()) == ("r1 extra", "ACGTN", "IIII#", b"RGZa\0")
    arr = ext.FastqRecordArrayView([v, v])
    assert len(arr) == 2 and arr[1].sequence() == "ACGTN" and arr.obj == v.obj * 2
    if _lib.device_count() <= 0:  # no CPU fallback: constructors that need the device fail loudly
        with pytest.raises(ext.SqGpuError, match="no CUDA device"):
            ext.QCMetrics()
