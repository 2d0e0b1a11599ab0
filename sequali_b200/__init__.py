"""sequali_b200 -- B200-native implementation of sequali's per-record QC hot path.

The public names are those of ``sequali/__init__.py`` in the reference, so
``import sequali_b200 as sequali`` is enough for callers of the hot path
(``src/sequali/__main__.py:24-42``).  Everything is implemented in
``sequali_b200._qc`` on top of ``libsqgpu.so`` (hand-written sm_100a kernels
behind the C ABI of ``include/sqgpu.h``).
"""
from ._qc import (  # noqa: F401
    A, C, G, N, T,
    AdapterCounter, BamParser, DedupEstimator, FastqParser, FastqRecordArrayView,
    FastqRecordView, InsertSizeMetrics, NanoStats, NanoStatsIterator, NanoporeReadInfo,
    OverrepresentedSequences, PerTileQuality, QCMetrics,
    DEFAULT_BASES_FROM_END, DEFAULT_BASES_FROM_START, DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS,
    DEFAULT_END_ANCHOR_LENGTH, DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH,
    DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET, DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH,
    DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET, DEFAULT_FRAGMENT_LENGTH,
    DEFAULT_MAX_UNIQUE_FRAGMENTS, DEFAULT_UNIQUE_SAMPLE_EVERY,
    INSERT_SIZE_MAX_ADAPTER_STORE_SIZE, MAX_SEQUENCE_SIZE, NUMBER_OF_NUCS, NUMBER_OF_PHREDS,
    PHRED_MAX, TABLE_SIZE,
)
from . import _qc  # noqa: F401

__version__ = "0.1.0"
