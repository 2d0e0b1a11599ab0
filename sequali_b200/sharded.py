"""Multi-GPU plumbing: one process per GPU, reads sharded contiguously.

The record stream is cut into contiguous shards (rank g owns global record
indices [lo_g, hi_g)); every rank runs the whole hot loop on its shard with its
own collectors; afterwards the tables are merged through a communicator (NCCL over
NVLink inside libsqgpu on the GPU box -- ``NcclComm``, no torch; a gloo stand-in in
the CPU tests).  No collective sits on the data path.

What merges exactly:
  * additive count tables -- QCMetrics, AdapterCounter, InsertSizeMetrics
    histogram, all counters: all-reduce(SUM) after padding to the longest
    ``max_length`` (all-reduce(MAX));
  * NanoStats: per-read records concatenated in rank order, min/max times by
    all-reduce;
  * PerTileQuality (``merge_pertile``): the per-(tile, position) double sums are
    chains in read order, so the lowest rank that holds a tile owns it; the
    other ranks hand their records of that tile over (whole records, in
    order) and the owner adds them behind its own reads -- bit-identical to one
    sequential pass.  With tile-sorted input only the tile at each shard
    border moves.
  * DedupEstimator (``merge_dedup``): the first rank owns the table; the
    others only hash, drop what fails the owner's final sampling mask (the
    mask only grows) and hand the rest over in read order.
  * OverrepresentedSequences (``merge_overrep``): the table travels rank by
    rank while it is not full (order matters); once full the key set is
    frozen, the remaining ranks count against it independently and the count
    arrays are summed.
The merge functions talk to the collectors through small adapters
(``Gpu*`` below for sequali_b200; the CPU tests plug a host-side stand-in in), so the
protocol itself is covered by world_size-2 gloo tests without a GPU.
"""
from __future__ import annotations

import os
import pickle

import numpy as np


def shard_bounds(n_records: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, near-equal shards: rank g owns [lo, hi)."""
    base, extra = divmod(n_records, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# ------------------------------------------------------------------------------
# the communicator the merges talk through
# ------------------------------------------------------------------------------
# Product: NcclComm (libsqgpu's sq_comm_*: NCCL on the context's launch stream, device buffers are
# DevBuf objects).  The CPU tests install a torch.distributed/gloo communicator of their own
# (tests/sharded_adapters.TorchComm) whose buffers are CPU tensors: the protocol below only uses
# `numel()`, `data_ptr()` and `zero_()` of a buffer, and hands buffers back to the communicator and
# to the collector adapter they came from.
class SoloComm:
    """world size 1: every collective is the identity."""
    rank, world = 0, 1

    def allreduce_host_u64(self, arr, op="sum"):
        return np.asarray(arr, dtype=np.uint64).copy()

    def bcast_bytes(self, data, src):
        return data

    def allgather_bytes(self, data):
        return [data]

    def barrier(self):
        pass

    def sync(self):
        pass


_COMM = SoloComm()


def use_comm(comm) -> None:
    """Install the communicator of this process (None: single rank)."""
    global _COMM
    _COMM = comm if comm is not None else SoloComm()


def comm():
    return _COMM


class DevBuf:
    """Stream-ordered device memory (sq_stream_alloc) with the three tensor methods the protocol uses."""

    def __init__(self, ctx, nbytes: int, itemsize: int = 1):
        self._ctx, self.nbytes, self.itemsize = ctx, int(nbytes), itemsize
        self.ptr = ctx.lib.sq_stream_alloc(ctx.h, max(self.nbytes, 16))
        if not self.ptr:
            from . import _lib
            raise MemoryError(_lib.last_error())

    def numel(self) -> int:
        return self.nbytes // self.itemsize

    def data_ptr(self) -> int:
        return self.ptr

    def zero_(self):
        self._ctx.lib.sq_stream_memset(self._ctx.h, self.ptr, 0, self.nbytes)
        return self

    def narrow(self, dim: int, start: int, length: int) -> "DevBuf":
        """A view of `length` items from item `start` (torch.Tensor.narrow's signature); the parent owns the memory."""
        view = object.__new__(DevBuf)
        view._ctx, view.itemsize, view.nbytes = self._ctx, self.itemsize, int(length) * self.itemsize
        view.ptr, view._parent = self.ptr + int(start) * self.itemsize, self
        return view

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr and getattr(self, "_parent", None) is None:
            try:
                self._ctx.lib.sq_stream_free(self._ctx.h, ptr)
            except Exception:
                pass


class NcclComm:
    """One process per GPU over NCCL, without torch: libsqgpu's sq_comm_* on the context's stream.

    The ncclUniqueId goes from rank 0 to the others through a file in /tmp named after the launcher
    (the parent process all local ranks share) and MASTER_PORT: ranks of one launch find each
    other, launches do not collide.  Single node, as the rest of the sharded path."""

    _serial = 0

    def __init__(self, rank: int, world: int, ctx=None, key: str | None = None):
        import ctypes as C
        import time
        from . import _lib
        self._ctx = ctx or _lib.Context.get()
        lib = self._ctx.lib
        self.rank, self.world = rank, world
        NcclComm._serial += 1
        key = key or f"{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}_{NcclComm._serial}"
        path = os.path.join(os.environ.get("SQ_COMM_DIR", "/tmp"), f"sqgpu_nccl_{key}.id")
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.check(lib.sq_comm_unique_id(ident), "sq_comm_unique_id")
            with open(path + ".tmp", "wb") as f:
                f.write(bytes(ident))
            os.replace(path + ".tmp", path)
        else:
            deadline = time.time() + 120
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise TimeoutError(f"rank 0 did not publish {path}")
                time.sleep(0.01)
            with open(path, "rb") as f:
                data = f.read()
            ident = (C.c_uint8 * 128).from_buffer_copy(data)
        h = C.c_void_p()
        _lib.check(lib.sq_comm_create(self._ctx.h, ident, rank, world, C.byref(h)), "sq_comm_create")
        self.h, self._path = h, path
        self.barrier()
        if rank == 0:
            try:
                os.remove(path)
            except OSError:
                pass

    @classmethod
    def from_env(cls, ctx=None):
        """RANK / WORLD_SIZE as torchrun (or any launcher) exports them; None for a single rank."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return None
        return cls(int(os.environ["RANK"]), world, ctx)

    def _check(self, rc, what):
        from . import _lib
        _lib.check(rc, what)

    # -- small host payloads ----------------------------------------------------------------------
    def allreduce_host_u64(self, arr, op="sum"):
        a = np.ascontiguousarray(arr, dtype=np.uint64).copy()
        code = {"sum": 0, "max": 1, "min": 2}[op]
        self._check(self._ctx.lib.sq_comm_allreduce_host_u64(self.h, a.ctypes.data, a.size, code),
                    "sq_comm_allreduce_host_u64")
        return a

    def bcast_bytes(self, data, src):
        n = self.allreduce_host_u64([len(data) if self.rank == src else 0], "max")[0]
        buf = np.zeros(max(int(n), 1), np.uint8)
        if self.rank == src:
            buf[:n] = np.frombuffer(data, np.uint8)
        self._check(self._ctx.lib.sq_comm_bcast_host(self.h, buf.ctypes.data, int(n), src), "sq_comm_bcast_host")
        return buf[:n].tobytes()

    def allgather_bytes(self, data):
        sizes = np.zeros(self.world, np.uint64)
        mine = np.array([len(data)], np.uint64)
        self._check(self._ctx.lib.sq_comm_allgather_host(self.h, mine.ctypes.data, sizes.ctypes.data, 8),
                    "sq_comm_allgather_host")
        width = int(sizes.max())
        if width == 0:
            return [b""] * self.world
        src = np.zeros(width, np.uint8)
        src[:len(data)] = np.frombuffer(data, np.uint8)
        out = np.zeros(width * self.world, np.uint8)
        self._check(self._ctx.lib.sq_comm_allgather_host(self.h, src.ctypes.data, out.ctypes.data, width),
                    "sq_comm_allgather_host")
        return [out[g * width:g * width + int(sizes[g])].tobytes() for g in range(self.world)]

    def barrier(self):
        self._check(self._ctx.lib.sq_comm_barrier(self.h), "sq_comm_barrier")

    def sync(self):
        # The collectives run on the stream the collectors' kernels run on, so collectors see received
        # buffers without this; the parser entry points work on their own stream, and received FASTQ
        # text (border tiles) goes through them: wait once per exchange.
        self._check(self._ctx.lib.sq_ctx_sync(self._ctx.h), "sq_ctx_sync")

    # -- device buffers ---------------------------------------------------------------------------
    def send(self, buf, dst):
        self._check(self._ctx.lib.sq_comm_send(self.h, buf.data_ptr(), buf.nbytes, dst), "sq_comm_send")

    def recv(self, buf, src):
        self._check(self._ctx.lib.sq_comm_recv(self.h, buf.data_ptr(), buf.nbytes, src), "sq_comm_recv")

    def bcast(self, buf, src):
        self._check(self._ctx.lib.sq_comm_bcast(self.h, buf.data_ptr(), buf.nbytes, src), "sq_comm_bcast")

    def allreduce_sum_u32(self, buf):
        self._check(self._ctx.lib.sq_comm_allreduce_u32(self.h, buf.data_ptr(), buf.numel(), 0),
                    "sq_comm_allreduce_u32")

    def group(self):
        comm_self = self

        class _Group:
            def __enter__(self):
                comm_self._check(comm_self._ctx.lib.sq_comm_group_start(comm_self.h), "sq_comm_group_start")

            def __exit__(self, *exc):
                comm_self._check(comm_self._ctx.lib.sq_comm_group_end(comm_self.h), "sq_comm_group_end")
                return False
        return _Group()

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self._ctx.lib.sq_comm_destroy(h)


def allreduce_max(value: int) -> int:
    return int(_COMM.allreduce_host_u64([int(value)], "max")[0])


def allreduce_min(value: int) -> int:
    return int(_COMM.allreduce_host_u64([int(value)], "min")[0])


def allreduce_sum_tables(tables: list[np.ndarray]) -> list[np.ndarray]:
    """One all-reduce(SUM) over a list of equally-shaped-per-rank u64 tables (host arrays)."""
    if _COMM.world == 1 or not tables:
        return [np.asarray(t, dtype=np.uint64) for t in tables]
    flat = np.concatenate([np.asarray(t, dtype=np.uint64).ravel() for t in tables])
    out = _COMM.allreduce_host_u64(flat, "sum")
    res, off = [], 0
    for src in tables:
        n = int(np.asarray(src).size)
        res.append(out[off:off + n].reshape(np.asarray(src).shape).copy())
        off += n
    return res


def _pad_rows(table: np.ndarray, rows: int, width: int, anchor_end: bool = False) -> np.ndarray:
    t = np.asarray(table, dtype=np.uint64).reshape(-1, width)
    if t.shape[0] == rows:
        return t
    out = np.zeros((rows, width), dtype=np.uint64)
    if anchor_end:
        out[rows - t.shape[0]:] = t
    else:
        out[:t.shape[0]] = t
    return out


def merge_qc(base, phred, ea_base, ea_phred, gc, phred_scores) -> dict:
    """QCMetrics tables of all ranks summed (per-position tables padded to the longest read)."""
    rows = allreduce_max(len(base) // 5)
    b, p = _pad_rows(base, rows, 5), _pad_rows(phred, rows, 12)
    merged = allreduce_sum_tables([b, p, np.asarray(ea_base, np.uint64), np.asarray(ea_phred, np.uint64),
                                   np.asarray(gc, np.uint64), np.asarray(phred_scores, np.uint64)])
    keys = ("base_count_table", "phred_count_table", "end_anchored_base_count_table",
            "end_anchored_phred_count_table", "gc_content", "phred_scores")
    return {k: v.ravel() for k, v in zip(keys, merged)}


def merge_adapter_counts(counts) -> list:
    """[(adapter, forward, reverse)] summed over ranks."""
    rows = allreduce_max(max([len(f) for _, f, _ in counts], default=0))
    padded = []
    for _, f, r in counts:
        padded += [_pad_rows(f, rows, 1).ravel(), _pad_rows(r, rows, 1).ravel()]
    merged = allreduce_sum_tables(padded)
    return [(a, merged[2 * i], merged[2 * i + 1]) for i, (a, _, _) in enumerate(counts)]


def gather_nanostats(infos: np.ndarray) -> np.ndarray:
    """Per-read NanoStats records of all ranks, concatenated in rank (= read) order."""
    if _COMM.world == 1:
        return infos
    parts = _COMM.allgather_bytes(infos.tobytes())
    return np.concatenate([np.frombuffer(p, dtype=infos.dtype) for p in parts])


def merge_tile_counts(tiles) -> tuple[list, list]:
    """PerTileQuality.get_tile_counts() of all ranks -> (merged, straddling tile ids).

    Tiles seen by a single rank are exact.  For a tile seen by several ranks the
    partial sums are added in rank order, which is not the reference's rounding
    order; such tiles are returned in ``straddling`` so that the caller can shard
    at tile borders instead."""
    if _COMM.world == 1:
        return list(tiles), []
    parts = _allgather_obj([(int(t), list(e), list(c)) for t, e, c in tiles])
    merged, straddling = {}, []
    for part in parts:  # rank order = read order
        for t, e, c in part:
            if t not in merged:
                merged[t] = (list(e), list(c))
                continue
            straddling.append(t)
            me, mc = merged[t]
            n = max(len(me), len(e))
            me += [0.0] * (n - len(me))
            mc += [0] * (n - len(mc))
            for i, (x, y) in enumerate(zip(e, c)):
                me[i] += x
                mc[i] += y
    return [(t, merged[t][0], merged[t][1]) for t in sorted(merged)], sorted(set(straddling))


# ------------------------------------------------------------------------------
# order-dependent collectors: exact merges
# ------------------------------------------------------------------------------
_NO_FAIL = (1 << 62)


def _rank_world():
    return _COMM.rank, _COMM.world


def _bcast_ints(values, src: int) -> list[int]:
    if _COMM.world == 1:
        return [int(v) for v in values]
    data = _COMM.bcast_bytes(np.array([int(v) for v in values], dtype=np.int64).tobytes(), src)
    return [int(v) for v in np.frombuffer(data, dtype=np.int64)]


def _allgather_obj(obj) -> list:
    if _COMM.world == 1:
        return [obj]
    return [pickle.loads(p) for p in _COMM.allgather_bytes(pickle.dumps(obj, protocol=4))]


def _allgather_u64_rows(block: np.ndarray) -> list[np.ndarray]:
    """All-gather of one 2-D uint64 block per rank (shapes may differ); blocks of all ranks in rank order."""
    block = np.ascontiguousarray(block, dtype=np.uint64)
    if _COMM.world == 1:
        return [block]
    shapes = _allgather_obj(tuple(block.shape))
    parts = _COMM.allgather_bytes(block.tobytes())
    return [np.frombuffer(p, dtype=np.uint64).reshape(sh).copy() for p, sh in zip(parts, shapes)]


def _bcast_u64(arr, src: int) -> np.ndarray:
    """Broadcast of a 1-D uint64 array from `src`."""
    if _COMM.world == 1:
        return np.asarray(arr, dtype=np.uint64)
    data = _COMM.bcast_bytes(np.ascontiguousarray(arr, dtype=np.uint64).tobytes() if _COMM.rank == src else b"", src)
    return np.frombuffer(data, dtype=np.uint64).copy()


def _comm_sync():
    """Make received buffers visible to the collectors (a no-op when the collectives share their stream)."""
    _COMM.sync()


def merge_dedup(dd) -> tuple[np.ndarray, dict]:
    """DedupEstimator over all shards.  ``dd``: rank 0 a live estimator that has seen its shard,
    other ranks deferred ones.  Adapter interface: ``modulo_bits()``, ``deferred_hashes(bits) ->
    buffer of int64`` (record order), ``consume(buffer)``, ``counts() -> np.ndarray``, ``info() ->
    dict``, ``empty(n) -> buffer``.  Returns (duplication counts in slot order, info) on every rank.

    The hashes of all the other ranks land in ONE buffer on the owner, in rank (= read) order, and
    are consumed in one go: what fails a later, higher sampling level is dropped by the table
    maintenance itself (the mask only grows, _qcmodule.c:4429-4431)."""
    rank, world = _rank_world()
    if world == 1:
        return dd.counts(), dd.info()
    bits = _bcast_ints([dd.modulo_bits() if rank == 0 else 0], 0)[0]
    mine = dd.deferred_hashes(bits) if rank > 0 else None
    sizes = [int(v) for v in _COMM.allreduce_host_u64(
        [0 if (mine is None or g != rank) else int(mine.numel()) for g in range(world)], "sum")]
    if rank == 0:
        total = sum(sizes[1:])
        if total:
            buf, at = dd.empty(total), 0
            with _COMM.group():
                for g in range(1, world):  # rank order = read order
                    if sizes[g]:
                        _COMM.recv(buf.narrow(0, at, sizes[g]), g)
                        at += sizes[g]
            _comm_sync()
            dd.consume(buf)
    elif sizes[rank]:
        with _COMM.group():
            _COMM.send(mine, 0)
    _comm_sync()
    info = dd.info() if rank == 0 else None
    keys = ("modulo_bits", "hash_table_size", "tracked_sequences")
    vals = _bcast_ints([info[k] for k in keys] if rank == 0 else [0] * len(keys), 0)
    counts = _bcast_u64(dd.counts() if rank == 0 else None, 0)
    return counts, dict(zip(keys, vals))


def merge_overrep(ov) -> None:
    """OverrepresentedSequences over all shards; afterwards every rank's collector holds the
    merged table and counters.  Adapter interface: ``state() -> (n_unique, full)``,
    ``table() -> (keys int64 tensor, counts int32 tensor)``, ``load(keys, counts | None,
    n_unique)``, ``apply_deferred()``, ``local_counters() -> [n_seqs, n_sampled, total_fragments,
    warn_records, first_warn]``, ``set_counters(...)``, ``empty_table()``."""
    rank, world = _rank_world()
    if world == 1:
        return
    cur = 0
    while True:
        n_unique, full = _bcast_ints(ov.state() if rank == cur else (0, 0), cur)
        if full or cur == world - 1:
            break
        if rank == cur:
            keys, counts = ov.table()
            with _COMM.group():
                _COMM.send(keys, cur + 1)
                _COMM.send(counts, cur + 1)
        elif rank == cur + 1:
            keys, counts = ov.empty_table()
            with _COMM.group():
                _COMM.recv(keys, cur)
                _COMM.recv(counts, cur)
            _comm_sync()
            ov.load(keys, counts, n_unique)
            ov.apply_deferred()
        cur += 1
    # `cur` holds the table of everything up to its own shard; behind it the key set is frozen
    keys, counts = ov.table() if rank == cur else ov.empty_table()
    _COMM.bcast(keys, cur)
    _comm_sync()
    if rank > cur:
        ov.load(keys, None, n_unique)
        ov.apply_deferred()
        _, counts = ov.table()
    elif rank < cur:
        counts.zero_()
    _COMM.allreduce_sum_u32(counts)  # wrap-around = the reference's u32 counter
    _comm_sync()
    ov.load(keys, counts, n_unique)
    local = ov.local_counters()  # reads, sampled reads, fragments, warnings of this shard only
    sums = allreduce_sum_tables([np.array(local[:4], dtype=np.uint64)])[0]
    first_warn = allreduce_min(int(local[4]) if local[4] >= 0 else _NO_FAIL)
    ov.set_counters(int(sums[0]), int(sums[1]), int(sums[2]), int(sums[3]),
                    -1 if first_warn == _NO_FAIL else first_warn)


def _tile_table(pt):
    """(ids int64[nt], sums float64[nt, w], counts uint64[nt, w]) of a collector: through its
    ``tile_table()`` when it has one (no Python lists on the way), else from ``tile_counts()``."""
    if hasattr(pt, "tile_table"):
        return pt.tile_table()
    tiles = pt.tile_counts()
    w = max([len(e) for _, e, _ in tiles], default=0)
    ids = np.array([int(t) for t, _, _ in tiles], dtype=np.int64)
    sums, counts = np.zeros((len(tiles), w), np.float64), np.zeros((len(tiles), w), np.uint64)
    for i, (_, e, c) in enumerate(tiles):
        sums[i, :len(e)] = e
        counts[i, :len(c)] = c
    return ids, sums, counts


def merge_pertile(pt, first_record: int) -> dict:
    """PerTileQuality over all shards -> {"tiles": [(tile, sums, counts)], "number_of_reads",
    "max_length", "skipped_record"} on every rank.  Adapter interface: ``tile_ids() -> list``,
    ``fail_index() -> local index of the first unparsable header or None``, ``number_of_reads()``,
    ``select(sorted tile ids, limit_records) -> list of uint8 tensors (FASTQ text, <= 2 GiB
    each)``, ``add_text(tensor)``, ``tile_counts()`` (optionally ``tile_table()``, see
    ``_tile_table``), ``empty(n) -> uint8 tensor``.  Tables travel as numpy arrays; the Python
    lists of the result are built once at the end."""
    rank, world = _rank_world()
    if world == 1:
        tiles = pt.tile_counts()
        return dict(tiles=tiles, number_of_reads=pt.number_of_reads(),
                    max_length=max([len(e) for _, e, _ in tiles], default=0), skipped_record=pt.fail_index())
    fail = pt.fail_index()
    fail_global = _NO_FAIL if fail is None else first_record + fail
    F = allreduce_min(fail_global)
    dropped = first_record >= F            # the whole shard lies behind the first unparsable header
    my_tiles = [] if dropped else sorted(int(t) for t in pt.tile_ids())
    my_reads = 0 if dropped else pt.number_of_reads()
    gathered = _allgather_obj((np.asarray(my_tiles, dtype=np.int64), my_reads))
    all_tiles = [g[0].tolist() for g in gathered]
    total_reads = int(sum(g[1] for g in gathered))
    owner = {}
    shared = False
    for g, ts in enumerate(all_tiles):
        for t in ts:
            if owner.setdefault(t, g) != g:
                shared = True
    if shared:  # some tile lives on two ranks: its records move to the owner, in read order
        outgoing = {}
        for o in range(rank):
            ids = [t for t in my_tiles if owner[t] == o]
            if ids:
                outgoing[o] = pt.select(ids, max(0, min(F - first_record, 1 << 62)))
        plan = _allgather_obj({o: [int(c.numel()) for c in chunks] for o, chunks in outgoing.items()})
        incoming = []  # (sender, buffer), in sender (= read) order
        with _COMM.group():
            for o in range(world):
                for g in range(o + 1, world):
                    for i, nbytes in enumerate(plan[g].get(o, [])):
                        if nbytes == 0:
                            continue
                        if rank == g:
                            _COMM.send(outgoing[o][i], o)
                        elif rank == o:
                            buf = pt.empty(nbytes)
                            _COMM.recv(buf, g)
                            incoming.append(buf)
        _comm_sync()
        for buf in incoming:
            pt.add_text(buf)
    # one uint64 block per rank: [tile id | sums (bit patterns) | counts] per owned tile
    if dropped:
        block = np.zeros((0, 1), np.uint64)
    else:
        ids, sums, counts = _tile_table(pt)
        keep = np.array([owner.get(int(t)) == rank for t in ids], dtype=bool)
        w = sums.shape[1]
        block = np.zeros((int(keep.sum()), 1 + 2 * w), np.uint64)
        block[:, 0] = ids[keep].astype(np.uint64)
        block[:, 1:1 + w] = np.ascontiguousarray(sums[keep], dtype=np.float64).view(np.uint64)
        block[:, 1 + w:] = counts[keep]
    parts = [p for p in _allgather_u64_rows(block) if p.shape[0]]
    width = max([(p.shape[1] - 1) // 2 for p in parts], default=0)
    n_all = sum(p.shape[0] for p in parts)
    ids = np.zeros(n_all, np.int64)
    sums, counts = np.zeros((n_all, width), np.float64), np.zeros((n_all, width), np.uint64)
    at = 0
    for p in parts:
        n, w = p.shape[0], (p.shape[1] - 1) // 2
        ids[at:at + n] = p[:, 0].astype(np.int64)
        sums[at:at + n, :w] = np.ascontiguousarray(p[:, 1:1 + w]).view(np.float64)
        counts[at:at + n, :w] = p[:, 1 + w:]
        at += n
    order = np.argsort(ids, kind="stable")
    tiles = list(zip(ids[order].tolist(), sums[order].tolist(), counts[order].tolist()))
    return dict(tiles=tiles, number_of_reads=total_reads, max_length=width,
                skipped_record=None if F == _NO_FAIL else F)


# ------------------------------------------------------------------------------
# adapters for the sequali_b200 collectors (device tensors, NCCL)
# ------------------------------------------------------------------------------
class GpuDedup:
    def __init__(self, estimator, deferred: bool):
        from ._lib import check
        self.dd, self._check = estimator, check
        if deferred:
            check(estimator._ctx.lib.sq_dedup_set_deferred(estimator._h, 1), "sq_dedup_set_deferred")

    def empty(self, n):
        return DevBuf(self.dd._ctx, int(n) * 8, 8)

    def modulo_bits(self):
        return self.dd._modulo_bits

    def deferred_hashes(self, bits):
        import ctypes as C
        self.dd._sync()
        n = C.c_uint64()
        lib = self.dd._ctx.lib
        self._check(lib.sq_dedup_deferred_compact(self.dd._h, bits, C.byref(n)), "sq_dedup_deferred_compact")
        _comm_sync()
        out = self.empty(n.value)
        self._check(lib.sq_dedup_deferred_fetch(self.dd._h, out.data_ptr()), "sq_dedup_deferred_fetch")
        return out

    def consume(self, t):
        self.dd._sync()
        self._check(self.dd._ctx.lib.sq_dedup_add_hashes(self.dd._h, t.data_ptr(), int(t.numel())),
                    "sq_dedup_add_hashes")

    def counts(self):
        return np.frombuffer(self.dd.duplication_counts(), dtype=np.uint64).copy()

    def info(self):
        i = self.dd._sync()
        return dict(modulo_bits=int(i.modulo_bits), hash_table_size=int(i.hash_table_size),
                    tracked_sequences=int(i.tracked_sequences))


class GpuOverrep:
    def __init__(self, collector, deferred: bool, first_record: int):
        from ._lib import check
        self.ov, self._check, self.first_record = collector, check, first_record
        check(collector._ctx.lib.sq_overrep_set_deferred(collector._h, 1 if deferred else 0, first_record),
              "sq_overrep_set_deferred")

    def _lib(self):
        return self.ov._ctx.lib

    def state(self):
        i = self.ov._sync()
        return int(i.collected_unique_fragments), int(i.collected_unique_fragments >= i.max_unique_fragments)

    def empty_table(self):
        size = int(self.ov._sync().table_size)
        return DevBuf(self.ov._ctx, size * 8, 8), DevBuf(self.ov._ctx, size * 4, 4)

    def table(self):
        keys, counts = self.empty_table()
        self._check(self._lib().sq_overrep_copy_table(self.ov._h, keys.data_ptr(), counts.data_ptr()),
                    "sq_overrep_copy_table")
        return keys, counts

    def load(self, keys, counts, n_unique):
        self.ov._sync()
        self._check(self._lib().sq_overrep_load_table(self.ov._h, keys.data_ptr(),
                                                      counts.data_ptr() if counts is not None else None,
                                                      n_unique), "sq_overrep_load_table")

    def apply_deferred(self):
        self._check(self._lib().sq_overrep_apply_deferred(self.ov._h), "sq_overrep_apply_deferred")
        self.ov._ctx.sync()

    def local_counters(self):
        i = self.ov._sync()
        first = int(i.first_warn_record) if i.warn_records else -1
        return [int(i.number_of_sequences) - self.first_record, int(i.sampled_sequences),
                int(i.total_fragments), int(i.warn_records), first]

    def set_counters(self, n_seqs, n_sampled, total_frags, warn_records, first_warn):
        self.ov._warned = max(self.ov._warned, warn_records)  # the shard that met it has warned already
        self._check(self._lib().sq_overrep_set_counters(
            self.ov._h, n_seqs, n_sampled, total_frags, warn_records,
            first_warn if first_warn >= 0 else 0xFFFFFFFFFFFFFFFF), "sq_overrep_set_counters")


class GpuPerTile:
    """Keeps the record arrays of the shard alive until the merge: records of tiles owned by a
    lower rank are copied out of them."""

    def __init__(self, collector):
        from ._lib import check
        self.pt, self._check, self.arrays, self._keep = collector, check, [], []

    def add_record_array(self, arr):
        self.arrays.append(arr)
        self.pt.add_record_array(arr)

    def empty(self, n):
        return DevBuf(self.pt._ctx, int(n), 1)  # (sq_stream_alloc pads: the parser's vector loads look past the text)

    def tile_ids(self):
        return self.pt._tile_arrays()[0].tolist()

    def tile_table(self):
        ids, err, cnt = self.pt._tile_arrays()
        return ids.astype(np.int64), err, cnt

    def fail_index(self):
        info = self.pt._sync()
        return int(info.skipped_record) if info.skipped else None

    def number_of_reads(self):
        return int(self.pt.number_of_reads)

    def tile_counts(self):
        return self.pt.get_tile_counts()

    def select(self, ids, limit_records):
        import ctypes as C
        lib = self.pt._ctx.lib
        self.pt._sync()
        idarr = np.asarray(sorted(ids), dtype=np.int64)
        idp = idarr.ctypes.data_as(C.c_void_p)
        sizes, base = [], 0
        for arr in self.arrays:
            n = C.c_uint64()
            lim = max(0, min(len(arr), limit_records - base))
            if lim:
                self._check(lib.sq_batch_select_tiles(arr._handle(), idp, len(idarr), lim, None, 0, C.byref(n)),
                            "sq_batch_select_tiles")
            sizes.append((lim, n.value))
            base += len(arr)
        # chunks of whole arrays' selections, at most 2 GiB each (a record array is < 4 GiB)
        chunks, cur, cur_bytes = [], [], 0
        for i, (lim, nb) in enumerate(sizes):
            if nb == 0:
                continue
            if cur and cur_bytes + nb > (1 << 31):
                chunks.append((cur, cur_bytes))
                cur, cur_bytes = [], 0
            cur.append(i)
            cur_bytes += nb
        if cur:
            chunks.append((cur, cur_bytes))
        out = []
        _comm_sync()
        for members, total in chunks:
            t = self.empty(total)
            off = 0
            for i in members:
                lim, nb = sizes[i]
                n = C.c_uint64()
                self._check(lib.sq_batch_select_tiles(self.arrays[i]._handle(), idp, len(idarr), lim,
                                                      t.data_ptr() + off, nb, C.byref(n)), "sq_batch_select_tiles")
                assert n.value == nb
                off += nb
            out.append(t)
        return out

    def add_text(self, t):
        import ctypes as C
        from . import _lib
        from ._qc import FastqRecordArrayView
        ctx = self.pt._ctx
        h, info = C.c_void_p(), _lib.ParseInfo()
        self._check(ctx.lib.sq_batch_from_device_fastq(ctx.h, t.data_ptr(), int(t.numel()), 2 ** 63, C.byref(h),
                                                       C.byref(info)), "sq_batch_from_device_fastq")
        assert info.consumed == t.numel()
        arr = FastqRecordArrayView._from_parser(h, info.n_records, None, int(t.numel()))
        self._keep.append((t, arr))  # the record array borrows the tensor's memory
        self.pt.add_record_array(arr)
        self.pt._sync()


# ------------------------------------------------------------------------------
# one rank's collectors for a contiguous shard of the read stream
# ------------------------------------------------------------------------------
class ShardedCollectors:
    """The single-end hot loop (src/sequali/__main__.py:279-306) on one shard, plus the merges.

    ``first_record`` is the global index of the shard's first read (it fixes which reads
    OverrepresentedSequences samples, _qcmodule.c:3833).  Rank 0 runs every collector as usual;
    on the other ranks DedupEstimator and OverrepresentedSequences only hash (their tables are
    order dependent and travel between ranks in ``merge``)."""

    def __init__(self, mod, adapters, first_record: int = 0, dedup_kwargs=None, overrep_kwargs=None):
        rank, _ = _rank_world()
        self.first_record = first_record
        self.qc, self.ns, self.ad = mod.QCMetrics(), mod.NanoStats(), mod.AdapterCounter(adapters)
        self.pt = GpuPerTile(mod.PerTileQuality())
        self.ov = GpuOverrep(mod.OverrepresentedSequences(**(overrep_kwargs or {})), rank > 0, first_record)
        self.dd = GpuDedup(mod.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=64,
                                                                       back_sequence_offset=0))), rank > 0)

    def add_record_array(self, arr) -> None:
        self.qc.add_record_array(arr)
        self.pt.add_record_array(arr)
        self.ov.ov.add_record_array(arr)
        self.ns.add_record_array(arr)
        self.ad.add_record_array(arr)
        self.dd.dd.add_record_array(arr)

    def merge(self) -> dict:
        """Merged results of all ranks, on every rank.  ``self.merge_ms`` afterwards: host wall time
        of each merge on this rank (waiting for slower ranks included).

        QCMetrics, AdapterCounter and NanoStats are merged in place on the device (sq_qc_allreduce,
        sq_adapters_allreduce, sq_nanostats_allgather): afterwards this rank's collectors answer the
        usual getters with the tables of the whole stream."""
        import time
        from . import _qc
        from ._lib import check
        t = [time.perf_counter()]
        self.merge_ms = {}

        def lap(name):
            t.append(time.perf_counter())
            self.merge_ms[name] = round((t[-1] - t[-2]) * 1e3, 2)

        c, qc = comm(), self.qc
        _qc._flush()
        if c.world > 1:
            lib = qc._ctx.lib
            check(lib.sq_qc_allreduce(qc._h, c.h), "sq_qc_allreduce")
            check(lib.sq_adapters_allreduce(self.ad._h, c.h), "sq_adapters_allreduce")
            check(lib.sq_nanostats_allgather(self.ns._h, c.h, self.first_record), "sq_nanostats_allgather")
        keys = ("base_count_table", "phred_count_table", "end_anchored_base_count_table",
                "end_anchored_phred_count_table", "gc_content", "phred_scores")
        out = dict(qc={k: np.frombuffer(getattr(qc, k)(), dtype=np.uint64) for k in keys})
        out["qc"]["number_of_reads"] = int(qc.number_of_reads)
        out["qc"]["max_length"] = int(qc.max_length)
        out["adapters"] = [(a, np.frombuffer(f, dtype=np.uint64), np.frombuffer(r, dtype=np.uint64))
                           for a, f, r in self.ad.get_counts()]
        out["adapters_number_of_sequences"] = int(self.ad.number_of_sequences)
        out["nano"] = self.ns  # holds the records of all ranks in read order
        lap("qc+adapters+nanostats (first getter: waits for the shard's kernels)")
        out["ptq"] = merge_pertile(self.pt, self.first_record)
        lap("pertile")
        counts, info = merge_dedup(self.dd)
        out["dedup"] = dict(counts=counts, **info)
        lap("dedup")
        merge_overrep(self.ov)
        out["overrep"] = self.ov.ov  # the merged table answers the usual getters
        lap("overrep")
        return out


# ------------------------------------------------------------------------------
# paired end (src/sequali/__main__.py:284-303): InsertSizeMetrics + the pair fingerprints
# ------------------------------------------------------------------------------
class GpuInsert:
    """InsertSizeMetrics of one shard; on ranks behind the first the adapter occurrences are kept for the
    owner of the two first-come tables (sq_insert_set_deferred)."""

    def __init__(self, collector, deferred: bool):
        from ._lib import check
        self.m, self._check = collector, check
        if deferred:
            check(collector._ctx.lib.sq_insert_set_deferred(collector._h, 1), "sq_insert_set_deferred")

    def kept(self, which: int):
        """(keys DevBuf of n x 32 bytes, hashes DevBuf of n x 8 bytes, n) in pair order."""
        import ctypes as C
        self.m._sync()
        lib, ctx = self.m._ctx.lib, self.m._ctx
        n = C.c_uint64()
        self._check(lib.sq_insert_deferred_count(self.m._h, which, C.byref(n)), "sq_insert_deferred_count")
        keys, hashes = DevBuf(ctx, n.value * 32, 32), DevBuf(ctx, n.value * 8, 8)
        if n.value:
            self._check(lib.sq_insert_deferred_fetch(self.m._h, which, keys.ptr, hashes.ptr), "sq_insert_deferred_fetch")
        return keys, hashes, n.value

    def add_keys(self, which: int, keys, hashes, n: int):
        self._check(self.m._ctx.lib.sq_insert_add_keys(self.m._h, which, keys.data_ptr(), hashes.data_ptr(), n),
                    "sq_insert_add_keys")


def merge_insert(ins: GpuInsert) -> dict:
    """InsertSizeMetrics over all shards, on every rank: the histogram and the counters are summed on the
    device; the occurrences of the other ranks reach the owner of the adapter tables in rank (= pair) order."""
    from ._lib import check
    rank, world = _rank_world()
    m = ins.m
    if world > 1:
        ctx = m._ctx
        for which in (0, 1):
            keys, hashes, n = ins.kept(which) if rank > 0 else (None, None, 0)
            sizes = [int(v) for v in _COMM.allreduce_host_u64([n if g == rank else 0 for g in range(world)], "sum")]
            total = sum(sizes[1:])
            if total == 0:
                continue
            if rank == 0:
                all_keys, all_hashes, at = DevBuf(ctx, total * 32, 32), DevBuf(ctx, total * 8, 8), 0
                with _COMM.group():
                    for g in range(1, world):
                        if sizes[g]:
                            _COMM.recv(all_keys.narrow(0, at, sizes[g]), g)
                            _COMM.recv(all_hashes.narrow(0, at, sizes[g]), g)
                            at += sizes[g]
                ins.add_keys(which, all_keys, all_hashes, total)
            elif n:
                with _COMM.group():
                    _COMM.send(keys, 0)
                    _COMM.send(hashes, 0)
        m._sync()
        check(ctx.lib.sq_insert_allreduce(m._h, _COMM.h), "sq_insert_allreduce")
    lists = _bcast_obj((m.adapters_read1(), m.adapters_read2()) if rank == 0 else None, 0)
    return dict(sizes=np.frombuffer(m.insert_sizes(), dtype=np.uint64), adapters1=lists[0], adapters2=lists[1],
                total_reads=int(m.total_reads), number_of_adapters_read1=int(m.number_of_adapters_read1),
                number_of_adapters_read2=int(m.number_of_adapters_read2))


def _bcast_obj(obj, src: int):
    if _COMM.world == 1:
        return obj
    data = _COMM.bcast_bytes(pickle.dumps(obj, protocol=4) if _COMM.rank == src else b"", src)
    return pickle.loads(data)


class ShardedPairedCollectors:
    """The paired-end hot loop on one shard of a pair of files, plus the merges: QCMetrics, PerTileQuality
    and OverrepresentedSequences per read side, DedupEstimator on the pair fingerprints, InsertSizeMetrics."""

    def __init__(self, mod, first_record: int = 0, dedup_kwargs=None, overrep_kwargs=None):
        rank, _ = _rank_world()
        self.first_record = first_record
        self.qc = (mod.QCMetrics(), mod.QCMetrics())
        self.pt = (GpuPerTile(mod.PerTileQuality()), GpuPerTile(mod.PerTileQuality()))
        self.ov = tuple(GpuOverrep(mod.OverrepresentedSequences(**(overrep_kwargs or {})), rank > 0, first_record)
                        for _ in range(2))
        self.dd = GpuDedup(mod.DedupEstimator(**(dedup_kwargs or dict(front_sequence_offset=0,
                                                                       back_sequence_offset=0))), rank > 0)
        self.ins = GpuInsert(mod.InsertSizeMetrics(), rank > 0)

    def add_record_array_pair(self, a, b) -> None:
        self.qc[0].add_record_array(a)
        self.pt[0].add_record_array(a)
        self.ov[0].ov.add_record_array(a)
        self.dd.dd.add_record_array_pair(a, b)
        self.ins.m.add_record_array_pair(a, b)
        self.qc[1].add_record_array(b)
        self.pt[1].add_record_array(b)
        self.ov[1].ov.add_record_array(b)

    def merge(self) -> dict:
        from . import _qc
        from ._lib import check
        c = comm()
        _qc._flush()
        out = {}
        keys = ("base_count_table", "phred_count_table", "end_anchored_base_count_table",
                "end_anchored_phred_count_table", "gc_content", "phred_scores")
        for i, qc in enumerate(self.qc):
            if c.world > 1:
                check(qc._ctx.lib.sq_qc_allreduce(qc._h, c.h), "sq_qc_allreduce")
            out[f"qc{i + 1}"] = {k: np.frombuffer(getattr(qc, k)(), dtype=np.uint64) for k in keys}
            out[f"qc{i + 1}"]["number_of_reads"] = int(qc.number_of_reads)
            out[f"qc{i + 1}"]["max_length"] = int(qc.max_length)
        for i in range(2):
            out[f"ptq{i + 1}"] = merge_pertile(self.pt[i], self.first_record)
            merge_overrep(self.ov[i])
            out[f"overrep{i + 1}"] = self.ov[i].ov
        counts, info = merge_dedup(self.dd)
        out["dedup"] = dict(counts=counts, **info)
        out["insert"] = merge_insert(self.ins)
        return out
