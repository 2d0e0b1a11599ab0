"""ctypes front-end of the CPU checker (oracle/sq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  Nothing under
``sequali_b200/`` imports this module.

The classes mirror the reference collectors (src/sequali/_qcmodule.c) but take
an explicit ``(buf, recs)`` record array: ``buf`` is a bytes-like object and
``recs`` a numpy array of ``REC_DTYPE`` (offsets into ``buf``), the equivalent
of the reference's FastqRecordArrayView (_qcmodule.c:575-579).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsqoracle.so")

REC_DTYPE = np.dtype([
    ("name_off", "<u8"), ("seq_off", "<u8"), ("qual_off", "<u8"), ("tags_off", "<u8"),
    ("name_len", "<u4"), ("seq_len", "<u4"), ("tags_len", "<u4"), ("pad", "<u4"),
    ("err_sum", "<f8"),
])
NANO_DTYPE = np.dtype([
    ("start_time", "<i8"), ("duration", "<f4"), ("channel_id", "<i4"),
    ("length", "<u4"), ("pad", "<u4"), ("cumulative_error_rate", "<f8"),
    ("parent_id_hash", "<u8"),
])
assert REC_DTYPE.itemsize == 56 and NANO_DTYPE.itemsize == 40

E_NO_AT, E_NO_PLUS, E_LEN, E_ASCII, E_PHRED = 1, 2, 3, 4, 5


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "sq_oracle.c")
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libsqoracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for name in ("orc_qc_new", "orc_ad_new", "orc_ptq_new", "orc_ov_new",
                     "orc_dd_new", "orc_ns_new", "orc_is_new"):
            getattr(_lib, name).restype = C.c_void_p
        for name in ("orc_parse_fastq", "orc_decode_bam"):
            getattr(_lib, name).restype = C.c_int64
        for name in ("orc_first_non_ascii", "orc_ov_add", "orc_ov_entries",
                     "orc_dd_counts", "orc_is_adapters", "orc_murmur3",
                     "orc_wang64", "orc_wang64_inverse", "orc_insert_size",
                     "orc_dd_fingerprint_hash"):
            getattr(_lib, name).restype = C.c_uint64
    return _lib


def _p(a):
    """pointer to a numpy array / bytes-like as void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(a)) if isinstance(a, bytearray) else C.c_char_p(a),
                      C.c_void_p)
    raise TypeError(type(a))


def as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8)
    return np.frombuffer(buf, dtype=np.uint8)


class FastqFormatError(ValueError):
    def __init__(self, code, pos):
        super().__init__(f"fastq format error code={code} at byte {pos}")
        self.code, self.pos = code, pos


def parse_fastq(buf, max_records=None):
    """-> (recs, consumed).  Complete records only (_qcmodule.c:1093-1171)."""
    b = as_u8(buf)
    n = b.size
    cap = (int(np.count_nonzero(b == 10)) // 4) + 1
    if max_records is None:
        max_records = cap
    recs = np.zeros(cap, dtype=REC_DTYPE)
    consumed, code, pos = C.c_uint64(), C.c_int(), C.c_uint64()
    k = lib().orc_parse_fastq(_p(b), C.c_uint64(n), C.c_uint64(max_records), _p(recs),
                              C.c_uint64(cap), C.byref(consumed), C.byref(code),
                              C.byref(pos))
    if k < 0:
        raise FastqFormatError(code.value, pos.value)
    return recs[:k].copy(), consumed.value


def first_non_ascii(buf) -> int:
    b = as_u8(buf)
    return lib().orc_first_non_ascii(_p(b), C.c_uint64(b.size))


def decode_bam(stream):
    """BAM alignment records (after the header) -> (packed, recs, consumed, skipped)."""
    b = as_u8(stream)
    n = b.size
    packed = np.zeros((n * 4 + 2) // 3 + 16, dtype=np.uint8)
    cap = n // 36 + 1
    recs = np.zeros(cap, dtype=REC_DTYPE)
    consumed, skipped, plen = C.c_uint64(), C.c_uint64(), C.c_uint64()
    k = lib().orc_decode_bam(_p(b), C.c_uint64(n), _p(packed), _p(recs), C.c_uint64(cap),
                             C.byref(consumed), C.byref(skipped), C.byref(plen))
    return packed[:plen.value].copy(), recs[:k].copy(), consumed.value, skipped.value


def pack_records(records):
    """[(name, seq, qual[, tags])] (str/bytes) -> (buf, recs) in the packed
    name|seq|qual|tags layout of FastqRecordArrayView.__new__ (_qcmodule.c:667-685)."""
    chunks, recs, off = [], np.zeros(len(records), dtype=REC_DTYPE), 0
    for i, rec in enumerate(records):
        name, seq, qual = (x.encode("ascii") if isinstance(x, str) else bytes(x)
                           for x in rec[:3])
        tags = bytes(rec[3]) if len(rec) > 3 and rec[3] is not None else b""
        assert len(seq) == len(qual)
        r = recs[i]
        r["name_off"], r["name_len"] = off, len(name)
        r["seq_off"], r["seq_len"] = off + len(name), len(seq)
        r["qual_off"] = off + len(name) + len(seq)
        r["tags_off"], r["tags_len"] = off + len(name) + 2 * len(seq), len(tags)
        chunks += [name, seq, qual, tags]
        off += len(name) + 2 * len(seq) + len(tags)
    return b"".join(chunks), recs


class _Handle:
    _free = None

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h and self._free:
            getattr(lib(), self._free)(C.c_void_p(h))


class PhredError(ValueError):
    pass


class QCMetrics(_Handle):
    _free = "orc_qc_free"

    def __init__(self, end_anchor_length=100):
        self.end_anchor_length = end_anchor_length
        self.h = lib().orc_qc_new(C.c_uint64(end_anchor_length))

    def add(self, buf, recs):
        """Also writes recs['err_sum'] (meta->accumulated_error_rate, :2126)."""
        b = as_u8(buf)
        bad, badrec = C.c_uint8(), C.c_uint64()
        rc = lib().orc_qc_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)),
                              C.byref(bad), C.byref(badrec))
        if rc:
            raise PhredError(f"Not a valid phred character: {chr(bad.value)}")

    def _info(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().orc_qc_info(C.c_void_p(self.h), C.byref(a), C.byref(b))
        return a.value, b.value

    max_length = property(lambda self: self._info()[0])
    number_of_reads = property(lambda self: self._info()[1])

    def tables(self):
        ml, ea = self.max_length, self.end_anchor_length
        t = dict(base=np.zeros(ml * 5, "<u8"), phred=np.zeros(ml * 12, "<u8"),
                 ea_base=np.zeros(ea * 5, "<u8"), ea_phred=np.zeros(ea * 12, "<u8"),
                 gc=np.zeros(101, "<u8"), mean_phred=np.zeros(94, "<u8"))
        lib().orc_qc_tables(C.c_void_p(self.h), *[_p(t[k]) for k in
                            ("base", "phred", "ea_base", "ea_phred", "gc", "mean_phred")])
        return t


class AdapterCounter(_Handle):
    _free = "orc_ad_free"

    def __init__(self, adapters):
        self.adapters = tuple(adapters)
        arr = (C.c_char_p * len(self.adapters))(*[a.encode("ascii") for a in self.adapters])
        self.h = lib().orc_ad_new(arr, C.c_uint64(len(self.adapters)))

    def add(self, buf, recs):
        b = as_u8(buf)
        lib().orc_ad_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)))

    def _info(self):
        a, b = C.c_uint64(), C.c_uint64()
        lib().orc_ad_info(C.c_void_p(self.h), C.byref(a), C.byref(b))
        return a.value, b.value

    max_length = property(lambda self: self._info()[0])
    number_of_sequences = property(lambda self: self._info()[1])

    def get_counts(self):
        out, ml = [], self.max_length
        for i, a in enumerate(self.adapters):
            f, r = np.zeros(ml, "<u8"), np.zeros(ml, "<u8")
            lib().orc_ad_counts(C.c_void_p(self.h), C.c_uint64(i), _p(f), _p(r))
            out.append((a, f, r))
        return out


class PerTileQuality(_Handle):
    _free = "orc_ptq_free"

    def __init__(self):
        self.h = lib().orc_ptq_new()
        self.skipped_record = None  # (call-relative) record whose header failed

    def add(self, buf, recs):
        b = as_u8(buf)
        skip, bad = C.c_uint64(), C.c_uint8()
        rc = lib().orc_ptq_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)),
                               C.byref(skip), C.byref(bad))
        if rc == 1:
            self.skipped_record = skip.value
        elif rc < 0:
            raise PhredError(f"Not a valid phred character: {chr(bad.value)}")
        return rc

    def _info(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        lib().orc_ptq_info(C.c_void_p(self.h), C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, bool(d.value)

    max_length = property(lambda self: self._info()[0])
    number_of_reads = property(lambda self: self._info()[1])
    skipped = property(lambda self: self._info()[3])

    def get_tile_counts(self):
        ml, _, nt, _ = self._info()
        ids = np.zeros(nt, "<u8")
        err = np.zeros(nt * ml, "<f8")
        cnt = np.zeros(nt * ml, "<u8")
        lib().orc_ptq_tables(C.c_void_p(self.h), _p(ids), _p(err), _p(cnt))
        return [(int(ids[i]), err[i * ml:(i + 1) * ml].copy(), cnt[i * ml:(i + 1) * ml].copy())
                for i in range(nt)]


def kmer_to_str(kmer: int, k: int) -> str:
    return "".join("ACGT"[(kmer >> (2 * (k - 1 - i))) & 3] for i in range(k))


class OverrepresentedSequences(_Handle):
    _free = "orc_ov_free"

    def __init__(self, max_unique_fragments=5_000_000, fragment_length=21, sample_every=8,
                 bases_from_start=100, bases_from_end=100):
        self.fragment_length = fragment_length
        self.h = lib().orc_ov_new(C.c_uint64(max_unique_fragments), C.c_uint64(fragment_length),
                                  C.c_uint64(sample_every), C.c_int64(bases_from_start),
                                  C.c_int64(bases_from_end))
        self.warned_records = 0

    def add(self, buf, recs):
        b = as_u8(buf)
        w = C.c_uint64()
        k = lib().orc_ov_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)),
                             C.byref(w))
        self.warned_records += k
        return k, w.value

    def info(self):
        a = np.zeros(6, "<u8")
        lib().orc_ov_info(C.c_void_p(self.h), _p(a))
        return dict(zip(("number_of_sequences", "sampled_sequences",
                         "collected_unique_fragments", "total_fragments", "table_size",
                         "max_unique_fragments"), map(int, a)))

    def entries(self):
        """(kmers u64, counts u32) of the stored fragments, slot order."""
        n = self.info()["collected_unique_fragments"]
        km, ct = np.zeros(n, "<u8"), np.zeros(n, "<u4")
        k = lib().orc_ov_entries(C.c_void_p(self.h), _p(km), _p(ct))
        assert k == n
        return km, ct

    def sequence_counts(self):
        km, ct = self.entries()
        return {kmer_to_str(int(a), self.fragment_length): int(c) for a, c in zip(km, ct)}


class DedupEstimator(_Handle):
    _free = "orc_dd_free"

    def __init__(self, max_stored_fingerprints=1_000_000, *, front_sequence_length=8,
                 back_sequence_length=8, front_sequence_offset=64, back_sequence_offset=64):
        self.h = lib().orc_dd_new(C.c_uint64(max_stored_fingerprints),
                                  C.c_uint64(front_sequence_length),
                                  C.c_uint64(back_sequence_length),
                                  C.c_uint64(front_sequence_offset),
                                  C.c_uint64(back_sequence_offset))

    def add(self, buf, recs):
        b = as_u8(buf)
        lib().orc_dd_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)))

    def add_pair(self, buf1, recs1, buf2, recs2):
        b1, b2 = as_u8(buf1), as_u8(buf2)
        assert len(recs1) == len(recs2)
        lib().orc_dd_add_pair(C.c_void_p(self.h), _p(b1), _p(recs1), _p(b2), _p(recs2),
                              C.c_uint64(len(recs1)))

    def add_raw_hash(self, h):
        lib().orc_dd_add_raw_hash(C.c_void_p(self.h), C.c_uint64(h))

    def info(self):
        a = np.zeros(3, "<u8")
        lib().orc_dd_info(C.c_void_p(self.h), _p(a))
        return dict(_modulo_bits=int(a[0]), _hash_table_size=int(a[1]),
                    tracked_sequences=int(a[2]))

    def duplication_counts(self, with_hashes=False):
        n = self.info()["tracked_sequences"]
        ct, hs = np.zeros(n, "<u8"), np.zeros(n, "<u8")
        k = lib().orc_dd_counts(C.c_void_p(self.h), _p(ct), _p(hs))
        assert k == n, (k, n)
        return (ct, hs) if with_hashes else ct


class NanoStats(_Handle):
    _free = "orc_ns_free"

    def __init__(self):
        self.h = lib().orc_ns_new()
        self.skipped_record = None

    def add(self, buf, recs):
        b = as_u8(buf)
        skip = C.c_uint64()
        rc = lib().orc_ns_add(C.c_void_p(self.h), _p(b), _p(recs), C.c_uint64(len(recs)),
                              C.byref(skip))
        if rc == 1:
            self.skipped_record = skip.value
        elif rc < 0:
            raise ValueError("malformed BAM tags")
        return rc

    def info(self):
        n, lo, hi, sk = C.c_uint64(), C.c_int64(), C.c_int64(), C.c_int()
        lib().orc_ns_info(C.c_void_p(self.h), C.byref(n), C.byref(lo), C.byref(hi), C.byref(sk))
        return dict(number_of_reads=n.value, minimum_time=lo.value, maximum_time=hi.value,
                    skipped=bool(sk.value))

    def infos(self):
        out = np.zeros(self.info()["number_of_reads"], NANO_DTYPE)
        lib().orc_ns_infos(C.c_void_p(self.h), _p(out))
        return out


class InsertSizeMetrics(_Handle):
    _free = "orc_is_free"

    def __init__(self, max_adapters=10_000):
        self.h = lib().orc_is_new(C.c_uint64(max_adapters))

    def add_pair(self, buf1, recs1, buf2, recs2):
        b1, b2 = as_u8(buf1), as_u8(buf2)
        assert len(recs1) == len(recs2)
        lib().orc_is_add_pair(C.c_void_p(self.h), _p(b1), _p(recs1), _p(b2), _p(recs2),
                              C.c_uint64(len(recs1)))

    def info(self):
        a = np.zeros(6, "<u8")
        lib().orc_is_info(C.c_void_p(self.h), _p(a))
        return dict(zip(("total_reads", "number_of_adapters_read1", "number_of_adapters_read2",
                         "max_insert_size", "entries_read1", "entries_read2"), map(int, a)))

    def insert_sizes(self):
        out = np.zeros(self.info()["max_insert_size"] + 1, "<u8")
        lib().orc_is_sizes(C.c_void_p(self.h), _p(out))
        return out

    def adapters(self, which):
        """[(adapter_str, count)] of read `which` (1 or 2) in slot order."""
        n = self.info()["entries_read%d" % which]
        seqs, cnt = np.zeros(n * 32, np.uint8), np.zeros(n, "<u8")
        k = lib().orc_is_adapters(C.c_void_p(self.h), C.c_int(which - 1), _p(seqs), _p(cnt))
        assert k == n
        raw = seqs.tobytes()
        return [(raw[i * 32 + 1: i * 32 + 1 + raw[i * 32]].decode("ascii"), int(cnt[i]))
                for i in range(n)]


def names_are_mates(n1: bytes, n2: bytes) -> bool:
    return bool(lib().orc_names_are_mates(C.c_char_p(n1), C.c_uint64(len(n1)),
                                          C.c_char_p(n2), C.c_uint64(len(n2))))


def error_table() -> np.ndarray:
    out = np.zeros(94, "<f8")
    lib().orc_error_table(_p(out))
    return out
