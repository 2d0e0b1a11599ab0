"""Adapters that let sequali_b200.sharded's merge protocol drive the CPU ORACLE
(test infrastructure): the same rank-to-rank hand-overs as on the GPUs, with
CPU tensors over gloo.  Used by tests/test_sharded.py only."""
import ctypes as C

import numpy as np
import torch

from oracle import oracle as orc


class TorchComm:
    """The communicator interface of sequali_b200.sharded over torch.distributed (gloo on CPU):
    buffers are CPU tensors.  Test infrastructure: the product's communicator is NcclComm."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce_host_u64(self, arr, op="sum"):
        a = np.ascontiguousarray(arr, dtype=np.uint64).copy()
        t = torch.from_numpy(a.view(np.int64))
        ops = {"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN}
        self.dist.all_reduce(t, op=ops[op])  # (values stay below 2^63 in the tests; sums wrap like u64)
        return a

    def bcast_bytes(self, data, src):
        box = [data]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]

    def allgather_bytes(self, data):
        parts = [None] * self.world
        self.dist.all_gather_object(parts, data)
        return parts

    def barrier(self):
        self.dist.barrier()

    def sync(self):
        pass

    def send(self, buf, dst):
        self.dist.send(buf, dst=dst)

    def recv(self, buf, src):
        self.dist.recv(buf, src=src)

    def bcast(self, buf, src):
        self.dist.broadcast(buf, src=src)

    def allreduce_sum_u32(self, buf):
        self.dist.all_reduce(buf)

    def group(self):
        import contextlib
        return contextlib.nullcontext()


def _tile_of(name: bytes):
    parts = name.split(b":")
    if len(parts) < 6 or not parts[4].isdigit() or not 1 <= len(parts[4]) <= 18:
        return -1
    return int(parts[4])


class OracleDedup:
    def __init__(self, deferred: bool, **kw):
        self.d = orc.DedupEstimator(**kw)
        self.deferred, self.hashes = deferred, []

    def add(self, buf, recs):
        if not self.deferred:
            self.d.add(buf, recs)
            return
        b = orc.as_u8(buf)
        fn = orc.lib().orc_dd_fingerprint_hash
        fn.restype = C.c_uint64
        for r in recs:
            seq = np.ascontiguousarray(b[int(r["seq_off"]):int(r["seq_off"]) + int(r["seq_len"])])
            self.hashes.append(fn(C.c_void_p(self.d.h), seq.ctypes.data_as(C.c_void_p), C.c_uint64(len(seq))))

    def empty(self, n):
        return torch.empty(int(n), dtype=torch.int64)

    def modulo_bits(self):
        return self.d.info()["_modulo_bits"]

    def deferred_hashes(self, bits):
        h = np.array(self.hashes, dtype=np.uint64)
        keep = h[(h & np.uint64((1 << bits) - 1)) == 0]
        return torch.from_numpy(keep.view(np.int64).copy())

    def consume(self, t):
        for h in t.numpy().view(np.uint64).tolist():
            self.d.add_raw_hash(h)

    def counts(self):
        return np.asarray(self.d.duplication_counts(), dtype=np.uint64)

    def info(self):
        i = self.d.info()
        return dict(modulo_bits=i["_modulo_bits"], hash_table_size=i["_hash_table_size"],
                    tracked_sequences=i["tracked_sequences"])


class OracleOverrep:
    def __init__(self, deferred: bool, first_record: int, **kw):
        self.o = orc.OverrepresentedSequences(**kw)
        self.deferred, self.first_record, self.kept = deferred, first_record, []
        self._set(first_record, 0, 0)

    def _set(self, n_seqs, n_sampled, total):
        orc.lib().orc_ov_set_counters(C.c_void_p(self.o.h), C.c_uint64(n_seqs), C.c_uint64(n_sampled),
                                      C.c_uint64(total))

    def add(self, buf, recs):
        if self.deferred:
            self.kept.append((buf, recs))
        else:
            self.o.add(buf, recs)

    def state(self):
        i = self.o.info()
        return i["collected_unique_fragments"], int(i["collected_unique_fragments"] >= i["max_unique_fragments"])

    def empty_table(self):
        size = self.o.info()["table_size"]
        return torch.empty(size, dtype=torch.int64), torch.empty(size, dtype=torch.int32)

    def table(self):
        keys, counts = self.empty_table()
        orc.lib().orc_ov_get_table(C.c_void_p(self.o.h), C.c_void_p(keys.data_ptr()), C.c_void_p(counts.data_ptr()))
        return keys, counts

    def load(self, keys, counts, n_unique):
        orc.lib().orc_ov_set_table(C.c_void_p(self.o.h), C.c_void_p(keys.data_ptr()),
                                   C.c_void_p(counts.data_ptr()) if counts is not None else None,
                                   C.c_uint64(n_unique))

    def apply_deferred(self):
        for buf, recs in self.kept:
            self.o.add(buf, recs)
        self.kept, self.deferred = [], False

    def local_counters(self):
        i = self.o.info()
        return [i["number_of_sequences"] - self.first_record, i["sampled_sequences"], i["total_fragments"], 0, -1]

    def set_counters(self, n_seqs, n_sampled, total_frags, warn_records, first_warn):
        self._set(n_seqs, n_sampled, total_frags)


class OraclePerTile:
    def __init__(self):
        self.p = orc.PerTileQuality()
        self.arrays, self.fail, self.n_seen = [], None, 0

    def add(self, buf, recs):
        self.arrays.append((bytes(buf), recs))
        rc = self.p.add(buf, recs)
        if rc == 1 and self.fail is None:
            self.fail = self.n_seen + self.p.skipped_record
        self.n_seen += len(recs)

    def empty(self, n):
        return torch.empty(int(n), dtype=torch.uint8)

    def tile_ids(self):
        return [t for t, _, _ in self.p.get_tile_counts()]

    def fail_index(self):
        return self.fail

    def number_of_reads(self):
        return self.p.number_of_reads

    def tile_counts(self):
        return [(t, e.tolist(), c.tolist()) for t, e, c in self.p.get_tile_counts()]

    def select(self, ids, limit_records):
        ids, out, base = set(ids), [], 0
        for buf, recs in self.arrays:
            for i, r in enumerate(recs):
                if base + i >= limit_records:
                    break
                no, nl = int(r["name_off"]), int(r["name_len"])
                if _tile_of(buf[no:no + nl]) in ids:
                    end = int(r["qual_off"]) + int(r["seq_len"]) + 1
                    out.append(buf[no - 1:end])
            base += len(recs)
        text = b"".join(out)
        return [torch.from_numpy(np.frombuffer(text, np.uint8).copy())] if text else []

    def add_text(self, t):
        text = t.numpy().tobytes()
        recs, consumed = orc.parse_fastq(text)
        assert consumed == len(text)
        self.p.add(text, recs)
