#!/usr/bin/env python
"""One or more all-module steps over HBM-resident synthetic reads, for ncu.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --reads 4000000 --steps 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--chunk-reads", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--total-reads", type=int, default=None,
                    help="length of the stream the reads are the head of (sets the reads per tile run)")
    args = ap.parse_args()
    import sequali_b200 as sq
    from sequali_b200 import _lib
    from sequali_b200.device import DeviceFastq
    ctx = _lib.Context.get()
    data = DeviceFastq.synth_illumina(args.reads, bench.READ_LENGTH, seed=2, chunk_reads=args.chunk_reads,
                                      total_reads=args.total_reads)
    for _ in range(args.steps):
        mods = bench.make_modules(sq)
        for arr in data.record_arrays():
            bench.feed(mods, arr)
        _, _, summary = bench.read_results(mods)
        ctx.sync()
    print(summary, ctx.launch_count)


if __name__ == "__main__":
    main()
