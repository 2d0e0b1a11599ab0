"""sequali_b200.ext -- the same public names as ``sequali_b200`` (and as the reference's
``sequali/__init__.py``), served by the CPython extension ``_qc`` (``qc_ext.cpp``) instead of the
ctypes mirror ``sequali_b200._qc``.  Build: ``make -C sequali_b200/ext`` (``__graft_entry__.build()``
does it).  Linking ``_qc.so`` into a directory named ``sequali`` next to the reference's unchanged
``__init__.py`` / ``util.py`` gives ``sequali._qc``."""
from ._qc import (  # noqa: F401
    A, C, G, N, T,
    AdapterCounter, BamParser, DedupEstimator, FastqParser, FastqRecordArrayView,
    FastqRecordView, InsertSizeMetrics, NanoStats, NanoStatsIterator, NanoporeReadInfo,
    OverrepresentedSequences, PerTileQuality, QCMetrics, SqGpuError,
    DEFAULT_BASES_FROM_END, DEFAULT_BASES_FROM_START, DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS,
    DEFAULT_END_ANCHOR_LENGTH, DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH,
    DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET, DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH,
    DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET, DEFAULT_FRAGMENT_LENGTH,
    DEFAULT_MAX_UNIQUE_FRAGMENTS, DEFAULT_UNIQUE_SAMPLE_EVERY,
    INSERT_SIZE_MAX_ADAPTER_STORE_SIZE, MAX_SEQUENCE_SIZE, NUMBER_OF_NUCS, NUMBER_OF_PHREDS,
    PHRED_MAX, TABLE_SIZE,
)
from . import _qc  # noqa: F401
