#!/usr/bin/env python
"""Where the time of one sharded step goes (run under torch.distributed.run, one rank per GPU)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--chunk-reads", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["SEQUALI_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sequali_b200 as sq
    from sequali_b200 import _lib, sharded
    from sequali_b200.device import DeviceFastq
    ctx = _lib.Context.get()
    data = DeviceFastq.synth_illumina(args.reads, bench.READ_LENGTH, seed=2, chunk_reads=args.chunk_reads,
                                      first_read=rank * args.reads, total_reads=args.reads * world)
    S = sharded
    acc = {}

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = round(acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3, 2)
            return r
        return w

    for name in ("_allgather_obj", "allreduce_max", "_bcast_obj", "_bcast_ints", "_tile_table", "allreduce_sum_tables"):
        setattr(S, name, timed(name, getattr(S, name)))
    # inside GpuPerTile.add_text: parse / add / sync
    from sequali_b200._qc import PerTileQuality
    PerTileQuality.add_record_array = timed("PerTileQuality.add_record_array", PerTileQuality.add_record_array)
    PerTileQuality._sync = timed("PerTileQuality._sync", PerTileQuality._sync)
    for cls, names in ((S.GpuPerTile, ("tile_ids", "fail_index", "number_of_reads", "select", "add_text")),
                       (S.GpuDedup, ("deferred_hashes", "consume", "counts")),
                       (S.GpuOverrep, ("table", "load", "apply_deferred"))):
        for name in names:
            setattr(cls, name, timed(cls.__name__ + "." + name, getattr(cls, name)))
    for step in range(args.steps):
        t = {}

        def lap(name, t0):
            ctx.sync()
            torch.cuda.synchronize()
            t[name] = round((time.perf_counter() - t0) * 1e3, 2)

        dist.barrier()
        acc.clear()
        t0 = time.perf_counter()
        coll = S.ShardedCollectors(sq, bench.ILLUMINA_ADAPTERS, first_record=rank * args.reads)
        for arr in data.record_arrays():
            coll.add_record_array(arr)
        lap("feed", t0)
        acc.clear()
        qc = coll.qc
        t0 = time.perf_counter()
        S.merge_qc(*[np.frombuffer(x, dtype=np.uint64) for x in (
            qc.base_count_table(), qc.phred_count_table(), qc.end_anchored_base_count_table(),
            qc.end_anchored_phred_count_table(), qc.gc_content(), qc.phred_scores())])
        S.merge_adapter_counts([(a, np.frombuffer(f, dtype=np.uint64), np.frombuffer(r, dtype=np.uint64))
                                for a, f, r in coll.ad.get_counts()])
        lap("qc+adapters", t0)
        t0 = time.perf_counter()
        S.merge_pertile(coll.pt, coll.first_record)
        lap("pertile", t0)
        t0 = time.perf_counter()
        S.merge_dedup(coll.dd)
        lap("dedup", t0)
        t0 = time.perf_counter()
        S.merge_overrep(coll.ov)
        coll.ov.ov.overrepresented_sequences(threshold_fraction=0.001, min_threshold=100)
        lap("overrep", t0)
        print(f"rank {rank} step {step}: {t} calls {acc}", flush=True)
        acc.clear()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
