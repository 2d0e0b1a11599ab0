"""Drop-in for the reference's ``sequali._qc`` extension, backed by libsqgpu.so.

Same type names, constructor arguments, method names, getters and error
behaviour as ``src/sequali/_qcmodule.c`` (stubs: ``src/sequali/_qc.pyi``); the
work itself runs in hand-written sm_100a kernels behind the C ABI of
``include/sqgpu.h``.  This module only moves bytes between Python objects and
that ABI and translates status structs into the reference's exceptions.

Execution is deferred: ``add_record_array`` enqueues kernels and returns;
getters and the reference's struct members (exposed here as properties)
synchronise first, which is the only point where the reference lets a caller
observe results.  ``add_read`` is synchronous, as tests expect
warnings and errors inside the call.

There is no CPU fallback: without libsqgpu.so or a CUDA device every
constructor raises ``SqGpuError``.
"""
from __future__ import annotations

import array
import ctypes as _C
import math
import sys
import warnings
from typing import Iterable, Optional

import numpy as np

from . import _lib
from ._lib import Context, check

# ---- module constants (reference _qcmodule.c:6082-6171) ----------------------
A, C, G, T, N = 0, 1, 2, 3, 4
NUMBER_OF_NUCS = 5
NUMBER_OF_PHREDS = 12
TABLE_SIZE = NUMBER_OF_NUCS * NUMBER_OF_PHREDS
PHRED_MAX = 93
MAX_SEQUENCE_SIZE = 64
DEFAULT_END_ANCHOR_LENGTH = 100
DEFAULT_MAX_UNIQUE_FRAGMENTS = 5_000_000
DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS = 1_000_000
DEFAULT_FRAGMENT_LENGTH = 21
DEFAULT_UNIQUE_SAMPLE_EVERY = 8
DEFAULT_BASES_FROM_START = 100
DEFAULT_BASES_FROM_END = 100
DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH = 8
DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH = 8
DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET = 64
DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET = 64
INSERT_SIZE_MAX_ADAPTER_STORE_SIZE = 31

# Staging size of the parsers when the caller does not pass one.  The reference
# defaults to 128 KiB (FASTQ) / 48 KiB (BAM) because its collectors work out of
# L1; a device launch wants tens of MiB, still small enough that the text of a
# batch stays resident in the 126 MB L2 between the collectors' kernels.
DEFAULT_FASTQ_BUFFERSIZE = 32 * 1024 * 1024
DEFAULT_BAM_BUFFERSIZE = 24 * 1024 * 1024

META_DTYPE = np.dtype([("name_off", "<u4"), ("name_len", "<u4"), ("seq_off", "<u4"),
                       ("seq_len", "<u4"), ("qual_off", "<u4"), ("tags_off", "<u4"),
                       ("tags_len", "<u4"), ("reserved", "<u4"), ("err_sum", "<f8")])
assert META_DTYPE.itemsize == 40

# 10^-(q/10): same values as the reference's generated table
# (score_to_error_rate.h), used only by FastqRecordView.__new__'s eager sum.
_ERROR_RATES = [10 ** -(q / 10) for q in range(PHRED_MAX + 1)]


def _void(arr: np.ndarray):
    return _C.c_void_p(arr.ctypes.data)


def _type_error(what: str, obj) -> TypeError:
    return TypeError(f"{what}, got {type(obj)!r}")


# ------------------------------------------------------------------------------
# record views
# ------------------------------------------------------------------------------
class FastqRecordView:
    """One record over a shared bytes buffer (reference :357-569)."""
    __slots__ = ("obj", "_meta")

    def __init__(self, name: str, sequence: str, qualities: str,
                 tags: Optional[bytes] = None):
        for label, value in (("name", name), ("sequence", sequence),
                             ("qualities", qualities)):
            if not isinstance(value, str):
                raise TypeError(f"FastqRecordView() argument '{label}' must be str, "
                                f"not {type(value).__name__}")
        if tags is not None and not isinstance(tags, bytes):
            raise TypeError("FastqRecordView() argument 'tags' must be bytes, "
                            f"not {type(tags).__name__}")
        try:
            name_b = name.encode("ascii")
        except UnicodeEncodeError:
            raise ValueError(f"name should contain only ASCII characters: {name!r}")
        try:
            seq_b = sequence.encode("ascii")
        except UnicodeEncodeError:
            raise ValueError(f"sequence should contain only ASCII characters: {sequence!r}")
        try:
            qual_b = qualities.encode("ascii")
        except UnicodeEncodeError:
            raise ValueError(f"qualities should contain only ASCII characters: {sequence!r}")
        if len(seq_b) != len(qual_b):
            raise ValueError("sequence and qualities have different lengths: "
                             f"{len(seq_b)} and {len(qual_b)}")
        tags_b = tags or b""
        total = len(name_b) + 2 * len(seq_b) + len(tags_b)
        if total > 0xFFFFFFFF:
            raise OverflowError("Total length of FASTQ record exceeds 4 GiB. "
                                f"Record name: {name!r}")
        err = 0.0
        rates = _ERROR_RATES
        for ch in qual_b:  # eager validation + plain left-to-right sum (:442-451)
            q = ch - 33
            if q < 0 or q > PHRED_MAX:
                raise ValueError(f"Not a valid phred character: {chr(ch)}")
            err += rates[q]
        self.obj = name_b + seq_b + qual_b + tags_b
        ln, ls = len(name_b), len(seq_b)
        self._meta = (0, ln, ln, ls, ln + ls, ln + 2 * ls, len(tags_b), err)

    @classmethod
    def _from_meta(cls, obj: bytes, meta) -> "FastqRecordView":
        self = object.__new__(cls)
        self.obj = obj
        self._meta = (int(meta["name_off"]), int(meta["name_len"]), int(meta["seq_off"]),
                      int(meta["seq_len"]), int(meta["qual_off"]), int(meta["tags_off"]),
                      int(meta["tags_len"]), float(meta["err_sum"]))
        return self

    def name(self) -> str:
        o, ln = self._meta[0], self._meta[1]
        return self.obj[o:o + ln].decode("ascii")

    def sequence(self) -> str:
        o, ln = self._meta[2], self._meta[3]
        return self.obj[o:o + ln].decode("ascii")

    def qualities(self) -> str:
        o, ln = self._meta[4], self._meta[3]
        return self.obj[o:o + ln].decode("ascii")

    def tags(self) -> bytes:
        o, ln = self._meta[5], self._meta[6]
        return self.obj[o:o + ln]

    def _packed(self):
        """(name|seq|qual|tags bytes, relative meta tuple) for array building."""
        no, nl, so, sl, qo, to, tl, err = self._meta
        o = self.obj
        return (o[no:no + nl] + o[so:so + sl] + o[qo:qo + sl] + o[to:to + tl], nl, sl, tl, err)


class _PinnedBuffer:
    """A pinned staging buffer.  cudaMallocHost is slow (it page-locks), so released buffers go back
    to a pool: requests are rounded up to size classes (powers of two below 1 MiB, multiples of 1 MiB
    above), the pool holds at most _POOL_BYTES in total and drops its least recently used blocks."""
    __slots__ = ("ptr", "size", "_cls", "_ctx")
    _pool: list = []      # [(class size, ptr)], oldest first
    _pool_bytes = 0
    _POOL_BYTES = 1 << 30

    @staticmethod
    def _size_class(n: int) -> int:
        n = max(n, 4096)
        if n >= 1 << 20:
            return (n + (1 << 20) - 1) & ~((1 << 20) - 1)
        return 1 << (n - 1).bit_length()

    def __init__(self, ctx: Context, size: int):
        self._ctx = ctx
        self.size = size
        self._cls = cls = self._size_class(size)
        pool = _PinnedBuffer._pool
        for i in range(len(pool) - 1, -1, -1):
            if pool[i][0] == cls:
                self.ptr = pool.pop(i)[1]
                _PinnedBuffer._pool_bytes -= cls
                return
        self.ptr = ctx.lib.sq_pinned_alloc(ctx.h, cls)
        if not self.ptr:
            raise MemoryError(_lib.last_error())

    def view(self, start: int = 0, stop: Optional[int] = None) -> memoryview:
        stop = self.size if stop is None else stop
        return memoryview((_C.c_char * (stop - start)).from_address(self.ptr + start)).cast("B")

    def __del__(self):
        ptr = getattr(self, "ptr", None)
        if not ptr:
            return
        self.ptr = None
        P = _PinnedBuffer
        P._pool.append((self._cls, ptr))
        P._pool_bytes += self._cls
        while P._pool_bytes > P._POOL_BYTES and P._pool:
            cls, old = P._pool.pop(0)
            P._pool_bytes -= cls
            self._ctx.lib.sq_pinned_free(self._ctx.h, old)


# ------------------------------------------------------------------------------
# deferred add_record_array
# ------------------------------------------------------------------------------
# Results of the collectors are observable only through getters and members
# (SURVEY.md 8b), so `add_record_array` may be deferred: the collectors fed with
# the same record array are gathered and handed to the device in ONE call
# (sq_fused_add), which walks the text once for all of them.  Anything that
# could observe state -- a getter, a member, add_read, another record array --
# flushes first; per collector the arrays are applied in call order.
_ROLES = ("qc", "pt", "ov", "ns", "ad", "dd")


class _Pending:
    array = None      # the record array being gathered (keeps it alive)
    mods: dict = {}   # role -> collector


def _flush() -> None:
    arr, mods = _Pending.array, _Pending.mods
    if arr is None:
        return
    _Pending.array, _Pending.mods = None, {}
    ctx = Context.get()
    hs = [mods[r]._h if r in mods else None for r in _ROLES]
    if "qc" in mods:
        arr._metas_stale = True
    check(ctx.lib.sq_fused_add(ctx.h, arr._handle(), *hs), "sq_fused_add")


_lib.FLUSH_HOOKS.append(_flush)  # Context.sync() drains the pending adds first


def _defer(role: str, mod, arr) -> None:
    if _Pending.array is not None:
        clash = (_Pending.array is not arr or role in _Pending.mods
                 # NanoStats copies the error sum QCMetrics left in the array (:5314): if it
                 # was fed before QCMetrics, keep that order
                 or (role == "qc" and "ns" in _Pending.mods))
        if clash:
            _flush()
    if _Pending.array is None:
        _Pending.array, _Pending.mods = arr, {}
    _Pending.mods[role] = mod


class FastqRecordArrayView:
    """A record array: bytes buffer + descriptors (reference :575-883), plus the
    handle of its device-resident copy."""

    def __init__(self, view_items: Iterable[FastqRecordView]):
        try:
            items = list(view_items)
        except TypeError:
            raise TypeError("view_items should be iterable")
        chunks, metas, off = [], np.zeros(len(items), dtype=META_DTYPE), 0
        for i, item in enumerate(items):
            if not isinstance(item, FastqRecordView):
                raise TypeError("Expected an iterable of FastqRecordView objects, but item "
                                f"{i} is of type {type(item)!r}: {item!r}")
            blob, nl, sl, tl, err = item._packed()
            metas[i] = (off, nl, off + nl, sl, off + nl + sl, off + nl + 2 * sl, tl, 0, err)
            chunks.append(blob)
            off += len(blob)
        self._init(b"".join(chunks), metas, None)

    # -- internal constructors ------------------------------------------------
    def _init(self, obj, metas, handle, n=None, pinned=None, nbytes=None):
        self._obj = obj            # bytes, or None while it only lives in `pinned`/device
        self._metas = metas        # numpy META_DTYPE array or None (still on device)
        self._h = handle           # sq_batch*
        self._n = len(metas) if n is None else n
        self._pinned = pinned      # (_PinnedBuffer, nbytes) backing a parser-made array
        self._nbytes = nbytes
        self._metas_stale = False  # err_sum changed on the device (QCMetrics ran)

    @classmethod
    def _from_parser(cls, handle, n, pinned, nbytes):
        self = object.__new__(cls)
        self._init(None, None, handle, n=n, pinned=pinned, nbytes=nbytes)
        return self

    @classmethod
    def _empty(cls):
        self = object.__new__(cls)
        self._init(b"", np.zeros(0, dtype=META_DTYPE), None)
        return self

    # -- device side ------------------------------------------------------------
    def _handle(self):
        """sq_batch* of this array, uploading a Python-built array on first use."""
        if self._h is None:
            ctx = Context.get()
            h = _C.c_void_p()
            buf = self._obj
            check(ctx.lib.sq_batch_from_packed(
                ctx.h, _C.cast(_C.c_char_p(buf), _C.c_void_p), len(buf),
                _void(self._metas), self._n, _C.byref(h)), "sq_batch_from_packed")
            self._h = h
        return self._h

    def _fetch_metas(self):
        if _Pending.array is self:
            _flush()
        if self._metas is None or self._metas_stale:
            m = np.zeros(self._n, dtype=META_DTYPE)
            if self._n:
                ctx = Context.get()
                check(ctx.lib.sq_batch_get_metas(self._h, _void(m)), "sq_batch_get_metas")
            self._metas = m
            self._metas_stale = False
        return self._metas

    @property
    def obj(self) -> bytes:
        if self._obj is None:
            if self._pinned is not None:
                buf, nbytes = self._pinned
                self._obj = bytes(buf.view(0, nbytes))
            else:
                out = np.empty(self._nbytes, dtype=np.uint8)
                ctx = Context.get()
                check(ctx.lib.sq_batch_get_bytes(self._h, _void(out)), "sq_batch_get_bytes")
                self._obj = out.tobytes()
        return self._obj

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i) -> FastqRecordView:
        n = self._n
        i = i.__index__()
        if i < 0:
            i += n
        if i < 0 or i >= n:
            raise IndexError("array index out of range")
        return FastqRecordView._from_meta(self.obj, self._fetch_metas()[i])

    def is_mate(self, other) -> bool:
        if not isinstance(other, FastqRecordArrayView):
            raise TypeError(f"other must be of type FastqRecordArrayView, got {type(other)!r}")
        if len(other) != len(self):
            raise ValueError("other is not the same length as this record array view. "
                             f"This length: {len(self)}, other length: {len(other)}")
        if self._n == 0:
            return True
        _flush()
        ctx = Context.get()
        first = _C.c_uint64()
        check(ctx.lib.sq_batch_is_mate(self._handle(), other._handle(), _C.byref(first)),
              "sq_batch_is_mate")
        return first.value == self._n

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None:
            self._h = None
            try:
                Context.get().lib.sq_batch_free(h)
            except Exception:
                pass


def _check_array(obj, name="record_array") -> FastqRecordArrayView:
    if not isinstance(obj, FastqRecordArrayView):
        raise TypeError(f"{name} should be a FastqRecordArrayView object, got {type(obj)!r}")
    return obj


def _check_read(obj) -> FastqRecordView:
    if not isinstance(obj, FastqRecordView):
        raise TypeError(f"read should be a FastqRecordView object, got {type(obj)!r}")
    return obj


def _single(read: FastqRecordView) -> FastqRecordArrayView:
    return FastqRecordArrayView([read])


# ------------------------------------------------------------------------------
# parsers
# ------------------------------------------------------------------------------
class FastqParser:
    """FastqParser(fileobj, initial_buffersize) -- reference :889-1244.

    The host side only stages bytes: ``readinto`` fills pinned memory, the
    record boundaries are found on the device (sq_batch_from_fastq).  The
    buffer-growing and leftover rules are those of
    FastqParser_create_record_array."""

    def __init__(self, fileobj, initial_buffersize: Optional[int] = None):
        explicit = initial_buffersize is not None
        size = initial_buffersize if explicit else DEFAULT_FASTQ_BUFFERSIZE
        if not isinstance(size, int):
            raise TypeError("initial_buffersize must be an integer")
        if size < 1:
            raise ValueError(f"initial_buffersize must be at least 1, got {size}")
        self._ctx = Context.get()
        self._file = fileobj
        self._read_in_size = size
        self._leftover = b""  # bytes after the last parsed record of the previous buffer

    def __iter__(self):
        return self

    def __next__(self) -> FastqRecordArrayView:
        arr = self._create_record_array(1, None)
        if len(arr) == 0:
            raise StopIteration
        return arr

    def read(self, number_of_records: int) -> FastqRecordArrayView:
        n = number_of_records.__index__()
        if n < 1:
            raise ValueError(f"number_of_records should be greater than 1, got {n}")
        return self._create_record_array(n, n)

    def _create_record_array(self, min_records: int, max_records: Optional[int]):
        ctx, lib = self._ctx, self._ctx.lib
        leftover = self._leftover
        step = self._read_in_size
        size = max(step, len(leftover) + (0 if len(leftover) < step else step))
        buf = _PinnedBuffer(ctx, size)
        buf.view(0, len(leftover))[:] = leftover
        filled = len(leftover)
        parsed = 0
        handle = None
        info = _lib.ParseInfo()
        readinto = self._file.readinto  # AttributeError for text files, like the reference
        while parsed < min_records:
            if filled == buf.size:  # grow by one step, keeping the content (:995-1020)
                bigger = _PinnedBuffer(ctx, buf.size + step)
                bigger.view(0, filled)[:] = buf.view(0, filled)
                buf = bigger
            got = readinto(buf.view(filled, buf.size))
            if got is None:
                got = 0
            new_filled = filled + got
            if new_filled == 0:
                break  # entire file is read
            if handle is not None:
                lib.sq_batch_free(handle)
                handle = None
            h = _C.c_void_p()
            rc = lib.sq_batch_from_fastq(ctx.h, buf.ptr, new_filled,
                                         max_records if max_records is not None else 2 ** 63,
                                         _C.byref(h), _C.byref(info))
            if rc == _lib.SQ_E_FORMAT:
                self._raise_format_error(buf, new_filled, info)
            check(rc, "sq_batch_from_fastq")
            handle = h
            parsed = info.n_records
            if got == 0:
                data = buf.view(0, new_filled)
                if bytes(data).count(b"\n") < 4:
                    lib.sq_batch_free(handle)
                    text = bytes(data).split(b"\0", 1)[0].decode("ascii", "replace")
                    raise EOFError(f"Incomplete record at the end of file {text}")
                filled = new_filled
                break
            filled = new_filled
        if handle is None:
            self._leftover = b""
            return FastqRecordArrayView._empty()
        consumed = info.consumed
        self._leftover = bytes(buf.view(consumed, filled))
        return FastqRecordArrayView._from_parser(handle, parsed, (buf, filled), filled)

    @staticmethod
    def _raise_format_error(buf, nbytes, info):
        data = buf.view(0, nbytes)
        code, pos = info.err_code, info.err_pos
        if code == _lib.PARSE_ASCII:
            raise ValueError("Found non-ASCII character in file: "
                             f"{bytes(data[pos:pos + 1]).decode('latin-1')}")
        if code == _lib.PARSE_NO_AT:
            raise ValueError(f"Record does not start with @ but with {chr(data[pos])}")
        if code == _lib.PARSE_NO_PLUS:
            raise ValueError("Record second header does not start with + but with "
                             f"{chr(data[pos])}")
        end = pos
        while end < nbytes and data[end] != 10:
            end += 1
        name = bytes(data[pos:end]).decode("ascii", "replace")
        raise ValueError(f"Record sequence and qualities do not have equal length, {name!r}")


class BamParser:
    """BamParser(fileobj, initial_buffersize) -- reference :1362-1725.

    The bytes read from the file are copied to the device once; the block_size
    chain, the drop of secondary / supplementary records and the 4-bit sequence /
    raw quality decode into the packed record array all run there
    (sq_batch_from_bam_bytes)."""

    def __init__(self, fileobj, initial_buffersize: Optional[int] = None):
        size = DEFAULT_BAM_BUFFERSIZE if initial_buffersize is None else initial_buffersize
        if size < 4:
            raise ValueError(f"initial_buffersize must be at least 4, got {size}")
        magic = fileobj.read(8)
        if type(magic) is not bytes:
            raise TypeError(f"file_obj {fileobj!r} is not a binary IO type, got {type(fileobj)!r}")
        if len(magic) < 8:
            raise EOFError("Truncated BAM file")
        if magic[:4] != b"BAM\1":
            raise ValueError(f"fileobj: {fileobj!r}, is not a BAM file. No BAM magic, "
                             f"instead found: {magic!r}")
        l_text = int.from_bytes(magic[4:8], "little")
        header = fileobj.read(l_text)
        if len(header) != l_text:
            raise EOFError("Truncated BAM file")
        n_ref_b = fileobj.read(4)
        if len(n_ref_b) != 4:
            raise EOFError("Truncated BAM file")
        self._n_ref = int.from_bytes(n_ref_b, "little")
        for _ in range(self._n_ref):
            l_name_b = fileobj.read(4)
            if len(l_name_b) != 4:
                raise EOFError("Truncated BAM file")
            chunk = int.from_bytes(l_name_b, "little") + 4
            if len(fileobj.read(chunk)) != chunk:
                raise EOFError("Truncated BAM file")
        self.header = header
        self._ctx = Context.get()
        self._file = fileobj
        self._read_in_size = size
        self._buf, self._filled = None, 0

    def __iter__(self):
        return self

    def __next__(self) -> FastqRecordArrayView:
        # [leftover | newly read bytes] live in one pinned buffer kept between calls
        ctx, lib = self._ctx, self._ctx.lib
        step = self._read_in_size
        while True:
            have = self._filled
            if have >= 4:
                want = max(int.from_bytes(self._buf.view(0, 4), "little"), step)  # :1527-1531
            else:
                want = step - have
            if self._buf is None or have + want > self._buf.size:
                bigger = _PinnedBuffer(ctx, max(have + want, (self._buf.size * 3) // 2 if self._buf else 0))
                if have:
                    _C.memmove(bigger.ptr, self._buf.ptr, have)
                self._buf = bigger
            got = self._file.readinto(self._buf.view(have, have + want)) or 0
            n = have + got
            if n == 0:
                raise StopIteration
            if got == 0:
                raise EOFError(f"Incomplete record at the end of file {bytes(self._buf.view(0, have))!r}")
            self._filled = n
            # record chain (:1623-1637) + decode of the bytes read so far, on the device
            h, plen = _C.c_void_p(), _C.c_uint64()
            kept, skipped, used = _C.c_uint64(), _C.c_uint64(), _C.c_uint64()
            check(lib.sq_batch_from_bam_bytes(ctx.h, self._buf.ptr, n, min(self._n_ref, 0x7fffffff), _C.byref(h),
                                              _C.byref(kept), _C.byref(skipped), _C.byref(used), _C.byref(plen)),
                  "sq_batch_from_bam_bytes")  # (synchronises)
            if kept.value or skipped.value:
                break
        pos, arr = used.value, None
        if kept.value:
            arr = FastqRecordArrayView._from_parser(h, kept.value, None, plen.value)
        left = self._filled - pos
        if left and pos:
            _C.memmove(self._buf.ptr, self._buf.ptr + pos, left)
        self._filled = left
        return arr if arr is not None else FastqRecordArrayView._empty()


# ------------------------------------------------------------------------------
# collectors
# ------------------------------------------------------------------------------
class _Collector:
    _destroy = None

    @staticmethod
    def _flush_pending():
        _flush()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and self._destroy:
            self._h = None
            try:
                getattr(Context.get().lib, self._destroy)(h)
            except Exception:
                pass


def _u64_array(np_arr: np.ndarray) -> array.array:
    out = array.array("Q")
    out.frombytes(np_arr.astype("<u8", copy=False).tobytes())
    return out


class QCMetrics(_Collector):
    """reference :1786-2385"""
    _destroy = "sq_qc_destroy"

    def __init__(self, end_anchor_length: int = DEFAULT_END_ANCHOR_LENGTH):
        ea = end_anchor_length.__index__()
        if ea < 0 or ea > 0xFFFFFFFF:
            raise ValueError(f"end_anchor_length must be between 0 and {0xFFFFFFFF}, got {ea}")
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_qc_create(self._ctx.h, ea, _C.byref(h)), "sq_qc_create")
        self._h = h

    def _sync(self) -> _lib.QcInfo:
        _flush()
        info = _lib.QcInfo()
        check(self._ctx.lib.sq_qc_sync(self._h, _C.byref(info)), "sq_qc_sync")
        if info.bad_phred:
            raise ValueError(f"Not a valid phred character: {chr(info.bad_phred_char)}")
        return info

    def add_read(self, read: FastqRecordView) -> None:
        _check_read(read)
        arr = _single(read)
        self.add_record_array(arr)
        self._sync()
        # QCMetrics stores the ordered error sum back into the record (:2126)
        read._meta = read._meta[:7] + (float(arr._fetch_metas()[0]["err_sum"]),)

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        arr = _check_array(record_array)
        if len(arr) == 0:
            return
        _defer("qc", self, arr)

    max_length = property(lambda self: self._sync().max_length)
    number_of_reads = property(lambda self: self._sync().number_of_reads)
    end_anchor_length = property(lambda self: self._sync().end_anchor_length)

    def _tables(self):
        info = self._sync()
        ml, ea = info.max_length, info.end_anchor_length
        t = [np.zeros(ml * 5, "<u8"), np.zeros(ml * 12, "<u8"), np.zeros(ea * 5, "<u8"),
             np.zeros(ea * 12, "<u8"), np.zeros(101, "<u8"), np.zeros(94, "<u8")]
        check(self._ctx.lib.sq_qc_read(self._h, *[_void(x) for x in t]), "sq_qc_read")
        return t

    def aggregate(self, data_ranges, count_thresholds=()):
        """Extension of the B200 build: what the report derives from base_count_table() and
        phred_count_table() for the given (start, stop) position ranges -- aggregate_count_matrix of both tables
        and the length distribution walk of SequenceLengthDistribution (report_modules.py:307-322, 575-637) --
        computed on the device tables; only the aggregated numbers come back.  See sequali_b200.report."""
        info = self._sync()
        ranges = np.asarray(list(data_ranges), dtype=np.uint64).reshape(-1, 2)
        starts, stops = np.ascontiguousarray(ranges[:, 0]), np.ascontiguousarray(ranges[:, 1])
        thr = np.asarray(list(count_thresholds), dtype=np.uint64)
        n = len(starts)
        base, phred = np.zeros(n * NUMBER_OF_NUCS, dtype=np.uint64), np.zeros(n * NUMBER_OF_PHREDS, dtype=np.uint64)
        lengths = np.zeros(n, dtype=np.uint64)
        summary = _lib.QcLengthSummary()
        check(self._ctx.lib.sq_qc_aggregate(self._h, _void(starts), _void(stops), n, _void(base), _void(phred),
                                            _void(lengths), _void(thr), len(thr), info.number_of_reads,
                                            _C.byref(summary)), "sq_qc_aggregate")
        return {"base_matrix": array.array("Q", base.tobytes()), "phred_matrix": array.array("Q", phred.tobytes()),
                "length_counts": lengths.tolist(), "total_bases": summary.total_bases,
                "minimum_length": summary.minimum_length, "n50": summary.n50, "n90": summary.n90,
                "threshold_lengths": list(summary.threshold_lengths[:len(thr)])}

    def base_count_table(self):
        return _u64_array(self._tables()[0])

    def phred_count_table(self):
        return _u64_array(self._tables()[1])

    def end_anchored_base_count_table(self):
        return _u64_array(self._tables()[2])

    def end_anchored_phred_count_table(self):
        return _u64_array(self._tables()[3])

    def gc_content(self):
        return _u64_array(self._tables()[4])

    def phred_scores(self):
        return _u64_array(self._tables()[5])


class AdapterCounter(_Collector):
    """reference :2406-2969"""
    _destroy = "sq_adapters_destroy"

    def __init__(self, adapters: Iterable[str]):
        adapters = tuple(adapters)  # TypeError "... not iterable" from Python itself
        if len(adapters) < 1:
            raise ValueError("At least one adapter is expected")
        for a in adapters:
            if type(a) is not str:
                raise TypeError("All adapter sequences must be of type str, "
                                f"got {type(a)!r}, for {a!r}")
            if not a.isascii():
                raise ValueError(f"Adapter must contain only ASCII characters: {a!r}")
            if len(a) > MAX_SEQUENCE_SIZE:
                raise ValueError(f"Maximum adapter size is {MAX_SEQUENCE_SIZE}, "
                                 f"got {len(a)} for {a!r}")
        self.adapters = adapters
        self._ctx = Context.get()
        arr = (_C.c_char_p * len(adapters))(*[a.encode("ascii") for a in adapters])
        h = _C.c_void_p()
        check(self._ctx.lib.sq_adapters_create(self._ctx.h, arr, len(adapters), _C.byref(h)),
              "sq_adapters_create")
        self._h = h

    def _sync(self):
        _flush()
        n, ml = _C.c_uint64(), _C.c_uint64()
        check(self._ctx.lib.sq_adapters_sync(self._h, _C.byref(n), _C.byref(ml)),
              "sq_adapters_sync")
        return n.value, ml.value

    number_of_sequences = property(lambda self: self._sync()[0])
    max_length = property(lambda self: self._sync()[1])

    def add_read(self, read: FastqRecordView) -> None:
        _check_read(read)
        self.add_record_array(_single(read))
        self._sync()

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        arr = _check_array(record_array)
        if len(arr):
            _defer("ad", self, arr)

    def get_counts(self):
        _, ml = self._sync()
        out = []
        for i, a in enumerate(self.adapters):
            f, r = np.zeros(ml, "<u8"), np.zeros(ml, "<u8")
            check(self._ctx.lib.sq_adapters_read(self._h, i, _void(f), _void(r)),
                  "sq_adapters_read")
            out.append((a, _u64_array(f), _u64_array(r)))
        return out


class PerTileQuality(_Collector):
    """reference :2975-3397"""
    _destroy = "sq_pertile_destroy"

    def __init__(self):
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_pertile_create(self._ctx.h, _C.byref(h)), "sq_pertile_create")
        self._h = h
        self._reason = None

    def _sync(self) -> _lib.PerTileInfo:
        _flush()
        info = _lib.PerTileInfo()
        check(self._ctx.lib.sq_pertile_sync(self._h, _C.byref(info)), "sq_pertile_sync")
        if info.bad_phred:
            raise ValueError(f"Not a valid phred character: {chr(info.bad_phred_char)}")
        if info.skipped and self._reason is None:
            buf = np.zeros(1 << 16, np.uint8)
            ln = _C.c_uint64()
            check(self._ctx.lib.sq_pertile_skipped_name(self._h, _void(buf), buf.size,
                                                        _C.byref(ln)), "sq_pertile_skipped_name")
            name = buf[:ln.value].tobytes().decode("ascii", "replace")
            self._reason = f"Can not parse header: {name!r}"
        return info

    max_length = property(lambda self: self._sync().max_length)
    number_of_reads = property(lambda self: self._sync().number_of_reads)

    @property
    def skipped_reason(self):
        self._sync()
        return self._reason

    def add_read(self, read: FastqRecordView) -> None:
        if self._reason is not None:
            return
        _check_read(read)
        self.add_record_array(_single(read))
        self._sync()

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        if self._reason is not None:
            return
        arr = _check_array(record_array)
        if len(arr):
            _defer("pt", self, arr)

    def _tile_arrays(self):
        """(tile ids u64[nt], sums f64[nt, max_length], counts u64[nt, max_length]) as numpy arrays."""
        info = self._sync()
        nt, ml = info.n_tiles, info.max_length
        ids, err, cnt = np.zeros(nt, "<u8"), np.zeros(nt * ml, "<f8"), np.zeros(nt * ml, "<u8")
        if nt:
            check(self._ctx.lib.sq_pertile_read(self._h, _void(ids), _void(err), _void(cnt)),
                  "sq_pertile_read")
        return ids, err.reshape(nt, ml), cnt.reshape(nt, ml)

    def get_tile_counts(self):
        ids, err, cnt = self._tile_arrays()
        if len(ids) == 0:
            return []
        ml = err.shape[1]
        if ml == 0:
            return [(t, [], []) for t in ids.tolist()]
        # counts[j] = reads of the tile longer than j: one number repeated for reads of one length,
        # so such a row is built from a single int object instead of max_length of them
        flat = (cnt == cnt[:, :1]).all(axis=1).tolist()
        first = cnt[:, 0].tolist()
        counts = [[first[i]] * ml if flat[i] else cnt[i].tolist() for i in range(len(first))]
        return list(zip(ids.tolist(), err.tolist(), counts))


def _kmer_to_sequence(kmer: int, k: int) -> str:
    return "".join("ACGT"[(kmer >> (2 * (k - 1 - i))) & 3] for i in range(k))


class OverrepresentedSequences(_Collector):
    """reference :3435-4236"""
    _destroy = "sq_overrep_destroy"

    def __init__(self, max_unique_fragments: int = DEFAULT_MAX_UNIQUE_FRAGMENTS,
                 fragment_length: int = DEFAULT_FRAGMENT_LENGTH,
                 sample_every: int = DEFAULT_UNIQUE_SAMPLE_EVERY,
                 bases_from_start: int = DEFAULT_BASES_FROM_START,
                 bases_from_end: int = DEFAULT_BASES_FROM_END):
        if max_unique_fragments < 1:
            raise ValueError("max_unique_fragments should be at least 1, got: "
                             f"{max_unique_fragments}")
        if (fragment_length & 1) == 0 or fragment_length > 31 or fragment_length < 3:
            raise ValueError("fragment_length must be between 3 and 31 and be an uneven "
                             f"number, got: {fragment_length}")
        if sample_every < 1:
            raise ValueError(f"sample_every must be 1 or greater. Got {sample_every}")
        self.max_unique_fragments = max_unique_fragments
        self.fragment_length = fragment_length
        self.sample_every = sample_every
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_overrep_create(self._ctx.h, max_unique_fragments, fragment_length,
                                              sample_every, bases_from_start, bases_from_end,
                                              _C.byref(h)), "sq_overrep_create")
        self._h = h
        self._warned = 0

    def _sync(self, source: Optional[FastqRecordArrayView] = None) -> _lib.OverrepInfo:
        _flush()
        info = _lib.OverrepInfo()
        check(self._ctx.lib.sq_overrep_sync(self._h, _C.byref(info)), "sq_overrep_sync")
        if info.warn_records > self._warned:
            self._warned = info.warn_records
            culprit = ""
            if source is not None and len(source) == 1:
                culprit = repr(source[0].sequence())
            warnings.warn("Sequence contains a chacter that is not A, C, G, T or N: "
                          f"{culprit}", UserWarning, stacklevel=3)
        return info

    number_of_sequences = property(lambda self: self._sync().number_of_sequences)
    sampled_sequences = property(lambda self: self._sync().sampled_sequences)
    collected_unique_fragments = property(
        lambda self: self._sync().collected_unique_fragments)
    total_fragments = property(lambda self: self._sync().total_fragments)

    def add_read(self, read: FastqRecordView) -> None:
        _check_read(read)
        arr = _single(read)
        self.add_record_array(arr)
        self._sync(arr)

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        arr = _check_array(record_array)
        if len(arr):
            _defer("ov", self, arr)

    def _entries(self):
        info = self._sync()
        n = info.collected_unique_fragments
        km, ct = np.zeros(n, "<u8"), np.zeros(n, "<u4")
        got = _C.c_uint64()
        check(self._ctx.lib.sq_overrep_read(self._h, _void(km), _void(ct), _C.byref(got)),
              "sq_overrep_read")
        return info, km[:got.value], ct[:got.value]

    def sequence_counts(self):
        _, km, ct = self._entries()
        k = self.fragment_length
        return {_kmer_to_sequence(int(a), k): int(c) for a, c in zip(km.tolist(), ct.tolist())}

    def overrepresented_sequences(self, threshold_fraction: float = 0.0001,
                                  min_threshold: int = 1,
                                  max_threshold: int = sys.maxsize):
        if threshold_fraction < 0.0 or threshold_fraction > 1.0:
            raise ValueError("threshold_fraction must be between 0.0 and 1.0 got, "
                             f"{threshold_fraction!r}")
        if min_threshold < 1:
            raise ValueError(f"min_threshold must be at least 1, got {min_threshold}")
        if max_threshold < 1:
            raise ValueError(f"max_threshold must be at least 1, got {max_threshold}")
        info = self._sync()
        sampled = info.sampled_sequences
        hits = math.ceil(threshold_fraction * sampled)
        hits = min(max_threshold, max(min_threshold, hits))
        k = self.fragment_length
        # the table is filtered on the device; only the hits cross PCIe
        cap = 4096
        while True:
            km, ct = np.zeros(cap, "<u8"), np.zeros(cap, "<u4")
            got = _C.c_uint64()
            check(self._ctx.lib.sq_overrep_read_min(self._h, min(hits, 0xFFFFFFFF), _void(km),
                                                    _void(ct), cap, _C.byref(got)),
                  "sq_overrep_read_min")
            if got.value <= cap:
                break
            cap = got.value
        km, ct = km[:got.value], ct[:got.value]
        result = [(int(c), int(c) / sampled, _kmer_to_sequence(int(a), k))
                  for a, c in zip(km.tolist(), ct.tolist())]
        result.sort(reverse=True)
        return result


class DedupEstimator(_Collector):
    """reference :4277-4802"""
    _destroy = "sq_dedup_destroy"

    def __init__(self, max_stored_fingerprints: int = DEFAULT_DEDUP_MAX_STORED_FINGERPRINTS, *,
                 front_sequence_length: int = DEFAULT_FINGERPRINT_FRONT_SEQUENCE_LENGTH,
                 back_sequence_length: int = DEFAULT_FINGERPRINT_BACK_SEQUENCE_LENGTH,
                 front_sequence_offset: int = DEFAULT_FINGERPRINT_FRONT_SEQUENCE_OFFSET,
                 back_sequence_offset: int = DEFAULT_FINGERPRINT_BACK_SEQUENCE_OFFSET):
        if max_stored_fingerprints < 100:
            raise ValueError("max_stored_fingerprints must be at least 100, not "
                             f"{max_stored_fingerprints}")
        for label, value in (("front_sequence_length", front_sequence_length),
                             ("back_sequence_length", back_sequence_length),
                             ("front_sequence_offset", front_sequence_offset),
                             ("back_sequence_offset", back_sequence_offset)):
            if value < 0:
                raise ValueError(f"{label} must be at least 0, got {value}.")
        if front_sequence_length + back_sequence_length == 0:
            raise ValueError("The sum of front_sequence_length and back_sequence_length "
                             "must be at least 0")
        self.front_sequence_length = front_sequence_length
        self.back_sequence_length = back_sequence_length
        self.front_sequence_offset = front_sequence_offset
        self.back_sequence_offset = back_sequence_offset
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_dedup_create(self._ctx.h, max_stored_fingerprints,
                                            front_sequence_length, back_sequence_length,
                                            front_sequence_offset, back_sequence_offset,
                                            _C.byref(h)), "sq_dedup_create")
        self._h = h

    def _sync(self) -> _lib.DedupInfo:
        _flush()
        info = _lib.DedupInfo()
        check(self._ctx.lib.sq_dedup_sync(self._h, _C.byref(info)), "sq_dedup_sync")
        return info

    _modulo_bits = property(lambda self: self._sync().modulo_bits)
    _hash_table_size = property(lambda self: self._sync().hash_table_size)
    tracked_sequences = property(lambda self: self._sync().tracked_sequences)

    @staticmethod
    def _seq_array(sequence: str, what="sequence") -> FastqRecordArrayView:
        if type(sequence) is not str:
            raise TypeError(f"sequence should be a str object, got {type(sequence)!r}")
        if not sequence.isascii():
            raise ValueError(f"{what} should consist only of ASCII characters.")
        arr = object.__new__(FastqRecordArrayView)
        metas = np.zeros(1, dtype=META_DTYPE)
        metas[0] = (0, 0, 0, len(sequence), 0, len(sequence), 0, 0, 0.0)
        arr._init(sequence.encode("ascii"), metas, None)
        return arr

    def add_sequence(self, sequence: str) -> None:
        self.add_record_array(self._seq_array(sequence))

    def add_sequence_pair(self, sequence1: str, sequence2: str) -> None:
        self.add_record_array_pair(self._seq_array(sequence1), self._seq_array(sequence2))

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        arr = _check_array(record_array)
        if len(arr):
            _defer("dd", self, arr)

    def add_record_array_pair(self, record_array1, record_array2) -> None:
        a1 = _check_array(record_array1, "record_array1")
        a2 = _check_array(record_array2, "record_array2")
        if len(a1) != len(a2):
            raise ValueError("record_array1 and record_array2 must be of the same size. "
                             f"Got {len(a1)} and {len(a2)} respectively.")
        if len(a1):
            _flush()
            check(self._ctx.lib.sq_dedup_add_pair(self._h, a1._handle(), a2._handle()),
                  "sq_dedup_add_pair")

    def duplication_counts(self):
        info = self._sync()
        # the counts land straight in the array the caller gets (no numpy detour: this getter
        # returns ~10^6 entries and its host time is part of every run)
        out = array.array("Q")
        out.frombytes(bytes(8 * info.tracked_sequences))  # (the constructor walks a bytes initialiser item by item)
        n = _C.c_uint64()
        check(self._ctx.lib.sq_dedup_read(self._h, _C.c_void_p(out.buffer_info()[0]), _C.byref(n)),
              "sq_dedup_read")
        if n.value < len(out):
            del out[n.value:]
        return out


class NanoporeReadInfo:
    """reference :4817-4880"""
    __slots__ = ("start_time", "channel_id", "length", "cumulative_error_rate", "duration",
                 "parent_id_hash")

    def __init__(self, rec):
        self.start_time = int(rec["start_time"])
        self.channel_id = int(rec["channel_id"])
        self.length = int(rec["length"])
        self.cumulative_error_rate = float(rec["cumulative_error_rate"])
        self.duration = float(rec["duration"])
        self.parent_id_hash = int(rec["parent_id_hash"])


NANO_DTYPE = np.dtype([("start_time", "<i8"), ("duration", "<f4"), ("channel_id", "<i4"),
                       ("length", "<u4"), ("reserved", "<u4"),
                       ("cumulative_error_rate", "<f8"), ("parent_id_hash", "<u8")])


class NanoStatsIterator:
    def __init__(self, infos: np.ndarray):
        self._infos, self._pos = infos, 0

    def __iter__(self):
        return self

    def __next__(self) -> NanoporeReadInfo:
        if self._pos == len(self._infos):
            raise StopIteration
        rec = self._infos[self._pos]
        self._pos += 1
        return NanoporeReadInfo(rec)


class NanoStats(_Collector):
    """reference :4882-5450"""
    _destroy = "sq_nanostats_destroy"

    def __init__(self):
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_nanostats_create(self._ctx.h, _C.byref(h)), "sq_nanostats_create")
        self._h = h
        self._reason = None
        self._pi_warned = 0

    def _sync(self) -> _lib.NanoStatsInfo:
        _flush()
        info = _lib.NanoStatsInfo()
        check(self._ctx.lib.sq_nanostats_sync(self._h, _C.byref(info)), "sq_nanostats_sync")
        if info.tag_error:  # the reference's exceptions for malformed aux data (:5078-5259)
            kind, detail = info.tag_error, info.tag_error_detail
            if kind == 2:
                raise ValueError(f"Invalid type for array {chr(detail & 0xff)}")
            if kind == 3:
                raise ValueError(f"Unknown tag type {chr(detail & 0xff)}")
            if kind == 4:
                tag, expected = (("st", "Z"), ("du", "f"), ("pi", "Z"))[min(detail >> 8, 2)]
                raise RuntimeError(f"Wrong tag type for '{tag}' expected '{expected}' got '{chr(detail & 0xff)}'")
            if kind == 5:
                raise SystemError("error return without exception set")
            raise ValueError("truncated tags")
        if info.pi_warnings > self._pi_warned:
            self._pi_warned = info.pi_warnings
            warnings.warn("pi tag should have a valid uuid4 format with 36 characters. "
                          f"Counted {info.pi_first_length}. Skipping tag.", UserWarning, stacklevel=3)
        if info.skipped and self._reason is None:
            buf = np.zeros(1 << 16, np.uint8)
            ln = _C.c_uint64()
            check(self._ctx.lib.sq_nanostats_skipped_name(self._h, _void(buf), buf.size,
                                                          _C.byref(ln)),
                  "sq_nanostats_skipped_name")
            name = buf[:ln.value].tobytes().decode("ascii", "replace")
            self._reason = f"Can not parse header: {name!r}"
        return info

    number_of_reads = property(lambda self: self._sync().number_of_reads)
    minimum_time = property(lambda self: self._sync().minimum_time)
    maximum_time = property(lambda self: self._sync().maximum_time)

    @property
    def skipped_reason(self):
        self._sync()
        return self._reason

    def add_read(self, read: FastqRecordView) -> None:
        _check_read(read)
        self.add_record_array(_single(read))
        self._sync()

    def add_record_array(self, record_array: FastqRecordArrayView) -> None:
        arr = _check_array(record_array)
        if self._reason is not None:
            return
        if len(arr):
            _defer("ns", self, arr)

    def report_tables(self, run_start_time: int, time_interval: int, time_slots: int) -> dict:
        """Extension of the B200 build: the per-read loop of NanoStatsReport.from_nanostats
        (report_modules.py:1990-2024) over the records on the device.  See sequali_b200.report."""
        self._sync()
        lib = self._ctx.lib
        n = time_slots
        t_bases, t_reads, t_active = (np.zeros(n, dtype=np.uint64) for _ in range(3))
        t_quals, speeds = np.zeros(n * 12, dtype=np.uint64), np.zeros(81, dtype=np.uint64)
        parents, n_ch, err = _C.c_uint64(), _C.c_uint64(), _lib.NanoReportError()
        check(lib.sq_nanostats_report(self._h, run_start_time, time_interval, n, _void(t_bases), _void(t_reads),
                                      _void(t_active), _void(t_quals), _void(speeds), _C.byref(parents),
                                      _C.byref(n_ch), _C.byref(err)), "sq_nanostats_report")
        ch = np.zeros(n_ch.value, dtype=np.int32)
        ch_bases, ch_err = np.zeros(n_ch.value, dtype=np.uint64), np.zeros(n_ch.value, dtype=np.float64)
        check(lib.sq_nanostats_report_channels(self._h, _void(ch), _void(ch_bases), _void(ch_err), n_ch.value),
              "sq_nanostats_report_channels")
        if err.kind == 1:
            raise OverflowError("cannot convert float infinity to integer")
        if err.kind == 2:
            raise ValueError("cannot convert float NaN to integer")
        if err.kind == 3:
            raise IndexError("list index out of range")
        return {"time_bases": t_bases.tolist(), "time_reads": t_reads.tolist(), "time_active_channels": t_active.tolist(),
                "time_qualities": t_quals.reshape(n, 12).tolist(), "translocation_speed": speeds.tolist(),
                "reads_with_parent": parents.value, "channels": ch.tolist(), "channel_bases": ch_bases.tolist(),
                "channel_cumulative_error": ch_err.tolist()}

    def nano_info_iterator(self) -> NanoStatsIterator:
        info = self._sync()
        out = np.zeros(info.number_of_reads, dtype=NANO_DTYPE)
        if info.number_of_reads:
            check(self._ctx.lib.sq_nanostats_read(self._h, _void(out)), "sq_nanostats_read")
        return NanoStatsIterator(out)


class InsertSizeMetrics(_Collector):
    """reference :5466-5982"""
    _destroy = "sq_insert_destroy"

    def __init__(self, max_adapters: int = 10_000):
        if max_adapters < 1:
            raise ValueError(f"max_adapters must be at least 1, got {max_adapters}")
        self._ctx = Context.get()
        h = _C.c_void_p()
        check(self._ctx.lib.sq_insert_create(self._ctx.h, max_adapters, _C.byref(h)),
              "sq_insert_create")
        self._h = h

    def _sync(self) -> _lib.InsertInfo:
        _flush()
        info = _lib.InsertInfo()
        check(self._ctx.lib.sq_insert_sync(self._h, _C.byref(info)), "sq_insert_sync")
        return info

    total_reads = property(lambda self: self._sync().total_reads)
    number_of_adapters_read1 = property(lambda self: self._sync().number_of_adapters_read1)
    number_of_adapters_read2 = property(lambda self: self._sync().number_of_adapters_read2)

    def add_sequence_pair(self, sequence1: str, sequence2: str) -> None:
        for i, s in enumerate((sequence1, sequence2), 1):
            if not isinstance(s, str):
                raise TypeError(f"InsertSizeMetrics.add_sequence_pair() argument {i} must be "
                                f"str, not {type(s).__name__}")
        a1 = DedupEstimator._seq_array(sequence1, "sequence1")
        a2 = DedupEstimator._seq_array(sequence2, "sequence2")
        self.add_record_array_pair(a1, a2)

    def add_record_array_pair(self, record_array1, record_array2) -> None:
        a1 = _check_array(record_array1, "record_array1")
        a2 = _check_array(record_array2, "record_array2")
        if len(a1) != len(a2):
            raise ValueError("record_array1 and record_array2 must be of the same size. "
                             f"Got {len(a1)} and {len(a2)} respectively.")
        if len(a1):
            _flush()
            check(self._ctx.lib.sq_insert_add_pair(self._h, a1._handle(), a2._handle()),
                  "sq_insert_add_pair")

    def insert_sizes(self):
        info = self._sync()
        out = np.zeros(info.max_insert_size + 1, "<u8")
        check(self._ctx.lib.sq_insert_read_sizes(self._h, _void(out)), "sq_insert_read_sizes")
        return _u64_array(out)

    def _adapters(self, which: int):
        info = self._sync()
        n = info.entries_read1 if which == 0 else info.entries_read2
        seqs, cnt = np.zeros(max(n, 1) * 32, np.uint8), np.zeros(max(n, 1), "<u8")
        got = _C.c_uint64()
        check(self._ctx.lib.sq_insert_read_adapters(self._h, which, _void(seqs), _void(cnt),
                                                    _C.byref(got)), "sq_insert_read_adapters")
        raw = seqs.tobytes()
        return [(raw[i * 32 + 1:i * 32 + 1 + raw[i * 32]].decode("ascii"), int(cnt[i]))
                for i in range(got.value)]

    def adapters_read1(self):
        return self._adapters(0)

    def adapters_read2(self):
        return self._adapters(1)
