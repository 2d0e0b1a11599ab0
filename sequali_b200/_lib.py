"""ctypes binding of libsqgpu.so (include/sqgpu.h).

There is no CPU fallback: importing this module works anywhere (so that the
package can be inspected and the symbol table tested on a CPU box), but the
first call that needs a device raises ``SqGpuError`` when the library or a
CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsqgpu.so")

SQ_OK, SQ_E_CUDA, SQ_E_ARG, SQ_E_NOMEM, SQ_E_FORMAT, SQ_E_NODEVICE, SQ_E_LIMIT = \
    0, -1, -2, -3, -4, -5, -6
PARSE_NO_AT, PARSE_NO_PLUS, PARSE_LEN, PARSE_ASCII = 1, 2, 3, 4


class SqGpuError(RuntimeError):
    pass


class Meta(C.Structure):
    _fields_ = [("name_off", C.c_uint32), ("name_len", C.c_uint32),
                ("seq_off", C.c_uint32), ("seq_len", C.c_uint32),
                ("qual_off", C.c_uint32), ("tags_off", C.c_uint32),
                ("tags_len", C.c_uint32), ("reserved", C.c_uint32),
                ("err_sum", C.c_double)]


class ParseInfo(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("consumed", C.c_uint64),
                ("n_newlines", C.c_uint64), ("max_seq_len", C.c_uint32),
                ("err_code", C.c_int32), ("err_record", C.c_uint64),
                ("err_pos", C.c_uint64)]


class QcInfo(C.Structure):
    _fields_ = [("number_of_reads", C.c_uint64), ("max_length", C.c_uint64),
                ("end_anchor_length", C.c_uint64), ("bad_phred", C.c_int32),
                ("bad_phred_char", C.c_uint8), ("bad_phred_record", C.c_uint64)]


class PerTileInfo(C.Structure):
    _fields_ = [("number_of_reads", C.c_uint64), ("max_length", C.c_uint64),
                ("n_tiles", C.c_uint64), ("skipped", C.c_int32),
                ("skipped_record", C.c_uint64), ("bad_phred", C.c_int32),
                ("bad_phred_char", C.c_uint8)]


class OverrepInfo(C.Structure):
    _fields_ = [("number_of_sequences", C.c_uint64), ("sampled_sequences", C.c_uint64),
                ("collected_unique_fragments", C.c_uint64), ("total_fragments", C.c_uint64),
                ("max_unique_fragments", C.c_uint64), ("table_size", C.c_uint64),
                ("warn_records", C.c_uint64), ("first_warn_record", C.c_uint64)]


class DedupInfo(C.Structure):
    _fields_ = [("modulo_bits", C.c_uint64), ("hash_table_size", C.c_uint64),
                ("tracked_sequences", C.c_uint64)]


class NanoInfo(C.Structure):
    _fields_ = [("start_time", C.c_int64), ("duration", C.c_float),
                ("channel_id", C.c_int32), ("length", C.c_uint32),
                ("reserved", C.c_uint32), ("cumulative_error_rate", C.c_double),
                ("parent_id_hash", C.c_uint64)]


class QcLengthSummary(C.Structure):
    _fields_ = [("total_bases", C.c_uint64), ("minimum_length", C.c_uint64), ("n50", C.c_uint64), ("n90", C.c_uint64),
                ("threshold_lengths", C.c_uint64 * 16)]


class NanoReportError(C.Structure):
    _fields_ = [("kind", C.c_int32), ("record", C.c_uint64)]


class NanoStatsInfo(C.Structure):
    _fields_ = [("number_of_reads", C.c_uint64), ("minimum_time", C.c_int64),
                ("maximum_time", C.c_int64), ("skipped", C.c_int32),
                ("skipped_record", C.c_uint64), ("tag_error", C.c_int32),
                ("tag_error_record", C.c_uint64), ("pi_warnings", C.c_uint64),
                ("tag_error_detail", C.c_uint32), ("pi_first_length", C.c_uint32)]


class BgzfBlock(C.Structure):
    _fields_ = [("comp_off", C.c_uint64), ("text_off", C.c_uint64), ("comp_len", C.c_uint32),
                ("text_len", C.c_uint32)]


class InsertInfo(C.Structure):
    _fields_ = [("total_reads", C.c_uint64), ("number_of_adapters_read1", C.c_uint64),
                ("number_of_adapters_read2", C.c_uint64), ("max_insert_size", C.c_uint64),
                ("entries_read1", C.c_uint64), ("entries_read2", C.c_uint64)]


assert C.sizeof(Meta) == 40 and C.sizeof(NanoInfo) == 40

_vp, _u64, _u32, _i64, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_int
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/sqgpu.h declares
SIGNATURES = {
    "sq_device_count": (_int, []),
    "sq_last_error": (C.c_char_p, []),
    "sq_device_numa_node": (_int, [_int]),
    "sq_ctx_create": (_int, [_int, _P(_vp)]),
    "sq_ctx_destroy": (None, [_vp]),
    "sq_ctx_block_cache_stats": (_int, [_vp, _P(_u64), _P(_u64), _P(_u64)]),
    "sq_ctx_sync": (_int, [_vp]),
    "sq_ctx_stream": (_vp, [_vp]),
    "sq_ctx_launch_count": (_u64, [_vp]),
    "sq_ctx_profile": (_int, [_vp, _int]),
    "sq_ctx_profile_report": (_int, [_vp, C.c_char_p, C.c_size_t]),
    "sq_timer_start": (_int, [_vp]),
    "sq_timer_stop": (_int, [_vp, _P(C.c_double)]),
    "sq_pinned_alloc": (_vp, [_vp, C.c_size_t]),
    "sq_pinned_free": (None, [_vp, _vp]),
    "sq_device_alloc": (_vp, [_vp, C.c_size_t]),
    "sq_device_free": (None, [_vp, _vp]),
    "sq_memcpy_h2d": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "sq_memcpy_d2h": (_int, [_vp, _vp, _vp, C.c_size_t]),
    "sq_batch_from_fastq": (_int, [_vp, _vp, _u64, _u64, _P(_vp), _P(ParseInfo)]),
    "sq_batch_from_device_fastq": (_int, [_vp, _vp, _u64, _u64, _P(_vp), _P(ParseInfo)]),
    "sq_batch_from_packed": (_int, [_vp, _vp, _u64, _vp, _u64, _P(_vp)]),
    "sq_bam_walk": (_int, [_vp, _u64, _vp, _u64, _P(_u64), _P(_u64), _P(_u64)]),
    "sq_batch_from_bam": (_int, [_vp, _vp, _u64, _vp, _u64, _P(_vp), _P(_u64)]),
    "sq_batch_from_bam_bytes": (_int, [_vp, _vp, _u64, C.c_int32, _P(_vp), _P(_u64), _P(_u64), _P(_u64), _P(_u64)]),
    "sq_qc_aggregate": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _u64, _u64, _vp]),
    "sq_nanostats_report": (_int, [_vp, C.c_int64, C.c_int64, _u64, _vp, _vp, _vp, _vp, _vp, _P(_u64), _P(_u64), _vp]),
    "sq_nanostats_report_channels": (_int, [_vp, _vp, _vp, _vp, _u64]),
    "sq_sequence_identity_batch": (_int, [_vp, _vp, _vp, _vp, _vp, _u64, _int, _int, _int, _int, _vp]),
    "sq_bam_walk_device": (_int, [_vp, _vp, _u64, C.c_int32, _vp, _u64, _P(_u64), _P(_u64), _P(_u64)]),
    "sq_fastq_stream_create": (_int, [_vp, _vp, _u64, _u64, _P(_vp)]),
    "sq_fastq_stream_next": (_int, [_vp, _P(_vp), _P(ParseInfo)]),
    "sq_fastq_stream_leftover": (_u64, [_vp]),
    "sq_fastq_stream_destroy": (None, [_vp]),
    "sq_bgzf_scan": (_int, [_vp, _u64, _vp, _u64, _P(_u64), _P(_u64), _P(_u64)]),
    "sq_bgzf_inflate": (_int, [_vp, _vp, _u64, _vp, _u64, _vp, _P(_u64), _P(_int)]),
    "sq_fastq_stream_create_bgzf": (_int, [_vp, _vp, _u64, _u64, _P(_vp)]),
    "sq_selftest_inflate_host": (_int, [_vp, _u32, _vp, _u32, _P(_u32)]),
    "sq_dedup_set_deferred": (_int, [_vp, _int]),
    "sq_dedup_deferred_compact": (_int, [_vp, _u64, _P(_u64)]),
    "sq_dedup_deferred_fetch": (_int, [_vp, _vp]),
    "sq_dedup_add_hashes": (_int, [_vp, _vp, _u64]),
    "sq_overrep_set_deferred": (_int, [_vp, _int, _u64]),
    "sq_overrep_apply_deferred": (_int, [_vp]),
    "sq_overrep_copy_table": (_int, [_vp, _vp, _vp]),
    "sq_overrep_load_table": (_int, [_vp, _vp, _vp, _u64]),
    "sq_overrep_set_counters": (_int, [_vp, _u64, _u64, _u64, _u64, _u64]),
    "sq_batch_select_tiles": (_int, [_vp, _vp, _u64, _u64, _vp, _u64, _P(_u64)]),
    "sq_batch_size": (_u64, [_vp]),
    "sq_batch_nbytes": (_u64, [_vp]),
    "sq_batch_max_seq_len": (_u32, [_vp]),
    "sq_batch_get_metas": (_int, [_vp, _vp]),
    "sq_batch_get_bytes": (_int, [_vp, _vp]),
    "sq_batch_is_mate": (_int, [_vp, _vp, _P(_u64)]),
    "sq_batch_free": (None, [_vp]),
    "sq_qc_create": (_int, [_vp, _u64, _P(_vp)]),
    "sq_qc_destroy": (None, [_vp]),
    "sq_qc_add": (_int, [_vp, _vp]),
    "sq_qc_sync": (_int, [_vp, _P(QcInfo)]),
    "sq_qc_read": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sq_adapters_create": (_int, [_vp, _P(C.c_char_p), _u64, _P(_vp)]),
    "sq_adapters_destroy": (None, [_vp]),
    "sq_adapters_add": (_int, [_vp, _vp]),
    "sq_adapters_sync": (_int, [_vp, _P(_u64), _P(_u64)]),
    "sq_adapters_read": (_int, [_vp, _u64, _vp, _vp]),
    "sq_pertile_create": (_int, [_vp, _P(_vp)]),
    "sq_pertile_destroy": (None, [_vp]),
    "sq_pertile_add": (_int, [_vp, _vp]),
    "sq_pertile_sync": (_int, [_vp, _P(PerTileInfo)]),
    "sq_pertile_skipped_name": (_int, [_vp, _vp, _u64, _P(_u64)]),
    "sq_pertile_read": (_int, [_vp, _vp, _vp, _vp]),
    "sq_overrep_create": (_int, [_vp, _u64, _u32, _u64, _i64, _i64, _P(_vp)]),
    "sq_overrep_destroy": (None, [_vp]),
    "sq_overrep_add": (_int, [_vp, _vp]),
    "sq_overrep_sync": (_int, [_vp, _P(OverrepInfo)]),
    "sq_overrep_read": (_int, [_vp, _vp, _vp, _P(_u64)]),
    "sq_overrep_read_min": (_int, [_vp, _u32, _vp, _vp, _u64, _P(_u64)]),
    "sq_dedup_create": (_int, [_vp, _u64, _u64, _u64, _u64, _u64, _P(_vp)]),
    "sq_dedup_destroy": (None, [_vp]),
    "sq_dedup_add": (_int, [_vp, _vp]),
    "sq_dedup_add_pair": (_int, [_vp, _vp, _vp]),
    "sq_dedup_sync": (_int, [_vp, _P(DedupInfo)]),
    "sq_dedup_read": (_int, [_vp, _vp, _P(_u64)]),
    "sq_nanostats_create": (_int, [_vp, _P(_vp)]),
    "sq_nanostats_destroy": (None, [_vp]),
    "sq_nanostats_add": (_int, [_vp, _vp]),
    "sq_nanostats_sync": (_int, [_vp, _P(NanoStatsInfo)]),
    "sq_nanostats_skipped_name": (_int, [_vp, _vp, _u64, _P(_u64)]),
    "sq_nanostats_read": (_int, [_vp, _vp]),
    "sq_insert_create": (_int, [_vp, _u64, _P(_vp)]),
    "sq_insert_destroy": (None, [_vp]),
    "sq_insert_add_pair": (_int, [_vp, _vp, _vp]),
    "sq_insert_sync": (_int, [_vp, _P(InsertInfo)]),
    "sq_insert_read_sizes": (_int, [_vp, _vp]),
    "sq_insert_read_adapters": (_int, [_vp, _int, _vp, _vp, _P(_u64)]),
    "sq_fused_add": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sq_comm_unique_id": (_int, [_vp]),
    "sq_comm_create": (_int, [_vp, _vp, _int, _int, _P(_vp)]),
    "sq_comm_destroy": (None, [_vp]),
    "sq_comm_rank": (_int, [_vp]),
    "sq_comm_world": (_int, [_vp]),
    "sq_comm_allreduce_u64": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_allreduce_u32": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_bcast": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_send": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_recv": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_group_start": (_int, [_vp]),
    "sq_comm_group_end": (_int, [_vp]),
    "sq_comm_allreduce_host_u64": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_bcast_host": (_int, [_vp, _vp, _u64, _int]),
    "sq_comm_allgather_host": (_int, [_vp, _vp, _vp, _u64]),
    "sq_comm_barrier": (_int, [_vp]),
    "sq_qc_allreduce": (_int, [_vp, _vp]),
    "sq_adapters_allreduce": (_int, [_vp, _vp]),
    "sq_nanostats_allgather": (_int, [_vp, _vp, _u64]),
    "sq_insert_set_deferred": (_int, [_vp, _int]),
    "sq_insert_deferred_count": (_int, [_vp, _int, _P(_u64)]),
    "sq_insert_deferred_fetch": (_int, [_vp, _int, _vp, _vp]),
    "sq_insert_add_keys": (_int, [_vp, _int, _vp, _vp, _u64]),
    "sq_insert_allreduce": (_int, [_vp, _vp]),
    "sq_stream_alloc": (_vp, [_vp, _u64]),
    "sq_stream_free": (None, [_vp, _vp]),
    "sq_stream_memset": (_int, [_vp, _vp, _int, _u64]),
    "sq_synth_illumina": (_int, [_vp, _vp, _u64, _u64, _u32, _u64, _u64, _u64, _P(_u64)]),
}

_lib = None
_lock = threading.Lock()
FLUSH_HOOKS: list = []  # callables run before Context.sync() (deferred adds of the _qc layer)


def load() -> C.CDLL:
    """dlopen libsqgpu.so (built in-tree by sequali_b200/csrc/Makefile)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SqGpuError(
                    f"{LIB_PATH} is missing: build it with `make -C sequali_b200/csrc` "
                    "(or __graft_entry__.build()); sequali_b200 has no CPU fallback")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here = ABI drift, fail loudly
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def device_count() -> int:
    return load().sq_device_count()


def last_error() -> str:
    return (load().sq_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc == SQ_OK:
        return
    msg = last_error()
    if rc == SQ_E_NOMEM:
        raise MemoryError(f"{what}: {msg}")
    if rc == SQ_E_ARG:
        raise ValueError(f"{what}: {msg}")
    raise SqGpuError(f"{what} failed (code {rc}): {msg}")


def prefetched(gen, depth: int = 1):
    """Advance the generator `gen` on a helper thread, `depth` items ahead of the consumer.

    The parsers' entry points of libsqgpu put their device work on the context's parser stream, so
    the record-boundary scan (and the host->device copy) of the NEXT record array overlaps with the
    collectors' kernels of the current one -- the read-ahead the reference gets from xopen's
    decompression threads (src/sequali/util.py:108-123).  Exceptions of `gen` surface at the item
    they belong to.  SEQUALI_B200_NO_PREFETCH=1 switches the helper thread off."""
    if os.environ.get("SEQUALI_B200_NO_PREFETCH"):
        yield from gen
        return
    import queue
    q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
    stop = threading.Event()

    def put(msg) -> bool:
        while not stop.is_set():
            try:
                q.put(msg, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def work():
        try:
            for item in gen:
                if not put(("item", item)):
                    return
            put(("done", None))
        except BaseException as e:  # noqa: BLE001 -- handed to the consumer
            put(("error", e))

    t = threading.Thread(target=work, name="sequali-b200-parser", daemon=True)
    t.start()
    try:
        while True:
            kind, val = q.get()
            if kind == "item":
                yield val
            elif kind == "error":
                raise val
            else:
                return
    finally:
        stop.set()
        t.join()
        while not q.empty():
            q.get_nowait()
        gen.close()


def bind_to_numa_node(lib, device: int):
    """Run this process on the cores of the NUMA node the GPU hangs off, so that the pinned staging
    buffers (first touch) and the threads that fill them sit next to the device: with eight ranks on a
    two-socket box, host->device copies from the far socket run at less than half speed.  Only when
    several ranks share the box (WORLD_SIZE > 1) or SEQUALI_B200_NUMA=1; SEQUALI_B200_NUMA=0 switches it
    off.  Returns the node, or None."""
    want = os.environ.get("SEQUALI_B200_NUMA")
    if want == "0" or (want is None and int(os.environ.get("WORLD_SIZE", "1")) <= 1):
        return None
    try:
        node = lib.sq_device_numa_node(device)
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError):
        pass
    return None


class Context:
    """One device context per process (device = $SEQUALI_B200_DEVICE, else
    $LOCAL_RANK, else 0)."""
    _instance = None

    def __init__(self, device: int | None = None):
        lib = load()
        if device is None:
            device = int(os.environ.get("SEQUALI_B200_DEVICE",
                                        os.environ.get("LOCAL_RANK", "0")))
        if lib.sq_device_count() <= 0:
            raise SqGpuError("no CUDA device visible: sequali_b200 needs a GPU "
                             "(there is no CPU fallback)")
        h = C.c_void_p()
        device = device % max(lib.sq_device_count(), 1)
        self.numa_node = bind_to_numa_node(lib, device)
        check(lib.sq_ctx_create(device, C.byref(h)), "sq_ctx_create")
        self.h, self.lib, self.device = h, lib, device

    @classmethod
    def get(cls) -> "Context":
        if cls._instance is None:
            cls._instance = Context()
        return cls._instance

    def sync(self):
        for hook in FLUSH_HOOKS:
            hook()
        check(self.lib.sq_ctx_sync(self.h), "sq_ctx_sync")

    @property
    def launch_count(self) -> int:
        return self.lib.sq_ctx_launch_count(self.h)

    def timer_start(self):
        check(self.lib.sq_timer_start(self.h), "sq_timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(self.lib.sq_timer_stop(self.h, C.byref(ms)), "sq_timer_stop")
        return ms.value

    def profile(self, enable: bool):
        check(self.lib.sq_ctx_profile(self.h, 1 if enable else 0), "sq_ctx_profile")

    def profile_report(self) -> dict:
        """{kernel: (launches, total_ms)} measured with CUDA events on the launch stream."""
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.sq_ctx_profile_report(self.h, buf, len(buf)), "sq_ctx_profile_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out
