"""HBM-resident FASTQ text (extension API, not part of the reference surface).

``DeviceFastq`` keeps uncompressed FASTQ text in device memory in
record-aligned chunks and hands each chunk to the collectors as a
``FastqRecordArrayView`` without any host copy -- the "inputs already resident
in HBM" leg of bench.py, and the natural entry point once decompression moves
to the device.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import Context, check
from ._qc import FastqRecordArrayView


class DeviceFastq:
    def __init__(self, ctx: Context):
        self._ctx = ctx
        self._base = None
        self.chunks: list[tuple[int, int, int]] = []  # (device pointer, bytes, reads)

    @property
    def nbytes(self) -> int:
        return sum(c[1] for c in self.chunks)

    @property
    def n_reads(self) -> int:
        return sum(c[2] for c in self.chunks)

    @classmethod
    def synth_illumina(cls, n_reads: int, read_length: int = 150, seed: int = 2,
                       chunk_reads: int = 1 << 18, first_read: int = 0,
                       total_reads: int | None = None) -> "DeviceFastq":
        """NovaSeq-style reads generated on the device (csrc/synth.cu, recipe C2)."""
        ctx = Context.get()
        self = cls(ctx)
        chunk_reads = min(chunk_reads, 1 << 23)  # 2.9 GB of text: u32 offsets reach 4 GiB
        n_chunks = (n_reads + chunk_reads - 1) // chunk_reads
        stride = (chunk_reads * (2 * read_length + 52) + 255) & ~255
        self._base = ctx.lib.sq_device_alloc(ctx.h, n_chunks * stride)
        if not self._base:
            raise MemoryError(_lib.last_error())
        # one run of reads per tile over the whole stream (all ranks' shards together), as in a
        # flow-cell-ordered file: a shard border cuts at most one tile
        total = total_reads or n_reads
        reads_per_tile = max(1, -(-total // 936))
        for c in range(n_chunks):
            n = min(chunk_reads, n_reads - c * chunk_reads)
            got = C.c_uint64()
            ptr = self._base + c * stride
            check(ctx.lib.sq_synth_illumina(ctx.h, ptr, stride, n, read_length, seed & 0xFFFF,
                                            first_read + c * chunk_reads, reads_per_tile,
                                            C.byref(got)), "sq_synth_illumina")
            self.chunks.append((ptr, got.value, n))
        ctx.sync()
        return self

    def record_arrays(self):
        """One FastqRecordArrayView per chunk; boundaries are found on the device, one chunk ahead of
        the consumer (parser stream + helper thread, see _lib.prefetched)."""
        return _lib.prefetched(self._record_arrays())

    def _record_arrays(self):
        ctx = self._ctx
        for ptr, nbytes, n in self.chunks:
            h, info = C.c_void_p(), _lib.ParseInfo()
            check(ctx.lib.sq_batch_from_device_fastq(ctx.h, ptr, nbytes, 2 ** 63, C.byref(h),
                                                     C.byref(info)), "sq_batch_from_device_fastq")
            assert info.n_records == n and info.consumed == nbytes, (info.n_records, n)
            yield FastqRecordArrayView._from_parser(h, n, None, nbytes)

    def to_host(self, max_reads: int | None = None):
        """(text, reads): concatenated text of the first chunks covering at least max_reads reads."""
        ctx = self._ctx
        take, reads = [], 0
        for ch in self.chunks:
            take.append(ch)
            reads += ch[2]
            if max_reads is not None and reads >= max_reads:
                break
        out = np.empty(sum(c[1] for c in take), dtype=np.uint8)
        off = 0
        for ptr, nbytes, _ in take:
            check(ctx.lib.sq_memcpy_d2h(ctx.h, out.ctypes.data + off, ptr, nbytes), "d2h")
            off += nbytes
        return out, reads

    def free(self):
        if self._base:
            self._ctx.lib.sq_device_free(self._ctx.h, self._base)
            self._base = None
            self.chunks = []

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HostFastq:
    """Uncompressed FASTQ text in PINNED host memory, fed to the device in windows.

    What a decompressor thread with a pinned output buffer hands over: every
    record array is one ``sq_batch_from_fastq`` call on a window of the host
    text (the H2D copy happens inside the call), the next window starts at the
    first byte the previous one did not consume.  No extra host copy."""

    def __init__(self, ctx: Context, nbytes: int):
        self._ctx = ctx
        self.nbytes = nbytes
        self.ptr = ctx.lib.sq_pinned_alloc(ctx.h, nbytes + 64)
        if not self.ptr:
            raise MemoryError(_lib.last_error())

    @classmethod
    def from_device(cls, data: "DeviceFastq", max_reads: int | None = None) -> tuple["HostFastq", int]:
        ctx = data._ctx
        take, reads = [], 0
        for ch in data.chunks:
            take.append(ch)
            reads += ch[2]
            if max_reads is not None and reads >= max_reads:
                break
        self = cls(ctx, sum(c[1] for c in take))
        off = 0
        for ptr, nbytes, _ in take:
            check(ctx.lib.sq_memcpy_d2h(ctx.h, self.ptr + off, ptr, nbytes), "d2h")
            off += nbytes
        return self, reads

    @classmethod
    def from_bytes(cls, data) -> "HostFastq":
        """Pinned copy of `data` (bytes-like)."""
        self = cls(Context.get(), len(data))
        self.view()[:] = data
        return self

    def view(self) -> memoryview:
        return memoryview((C.c_char * self.nbytes).from_address(self.ptr)).cast("B")

    def record_arrays(self, window: int = 256 << 20):
        """Record arrays of ~`window` bytes of text each (sq_fastq_stream_*: the host->device
        copies run ahead of the parser on their own stream, the parser one array ahead of the
        collectors)."""
        return _lib.prefetched(self._record_arrays(window))

    def record_arrays_bgzf(self, window: int = 256 << 20):
        """The same for BGZF-compressed FASTQ in this buffer (bgzip output): the members travel
        compressed and are inflated on the device (sq_fastq_stream_create_bgzf), `window` bytes of
        text per record array."""
        return _lib.prefetched(self._record_arrays(window, bgzf=True))

    def _record_arrays(self, window: int, bgzf: bool = False):
        ctx = self._ctx
        stream = C.c_void_p()
        create = ctx.lib.sq_fastq_stream_create_bgzf if bgzf else ctx.lib.sq_fastq_stream_create
        rc = create(ctx.h, self.ptr, self.nbytes, window, C.byref(stream))
        if rc == _lib.SQ_E_FORMAT:
            raise ValueError(_lib.last_error())
        check(rc, "sq_fastq_stream_create")
        try:
            while True:
                h, info = C.c_void_p(), _lib.ParseInfo()
                rc = ctx.lib.sq_fastq_stream_next(stream, C.byref(h), C.byref(info))
                if rc == _lib.SQ_E_FORMAT and info.err_code == 5:
                    raise ValueError(_lib.last_error())
                if rc == _lib.SQ_E_FORMAT:
                    raise ValueError(f"malformed FASTQ text: parse error {info.err_code} in record "
                                     f"{info.err_record} at byte {info.err_pos} of its record array")
                check(rc, "sq_fastq_stream_next")
                if not h:
                    break
                yield FastqRecordArrayView._from_parser(h, info.n_records, None, ctx.lib.sq_batch_nbytes(h))
            left = ctx.lib.sq_fastq_stream_leftover(stream)
            if left:
                raise EOFError(f"Incomplete record at the end of the text: {left} bytes")
        finally:
            ctx.lib.sq_fastq_stream_destroy(stream)

    def free(self):
        if self.ptr:
            self._ctx.lib.sq_pinned_free(self._ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
