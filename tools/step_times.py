#!/usr/bin/env python
"""Per-step wall times of the HBM-resident bench step (100 M reads by default), to see the spread between steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import sequali_b200 as sq
    from sequali_b200 import _lib
    from sequali_b200.device import DeviceFastq
    n = int(os.environ.get("N_READS", "100000000"))
    steps = int(os.environ.get("STEPS", "12"))
    ctx = _lib.Context.get()
    data = DeviceFastq.synth_illumina(n, bench.READ_LENGTH, seed=2, chunk_reads=1 << 22)
    times = []
    for _ in range(steps):
        ctx.sync()
        t0 = time.perf_counter()
        mods = bench.make_modules(sq)
        for arr in data.record_arrays():
            bench.feed(mods, arr)
        bench.read_results(mods)
        ctx.sync()
        times.append((time.perf_counter() - t0) * 1e3)
    print(" ".join(f"{t:.1f}" for t in times))
    if os.environ.get("SQ_TRACE_ALLOC"):
        del mods, data
        ctx.lib.sq_ctx_destroy(ctx.h)  # (prints the pool's high water marks; the process ends here)


if __name__ == "__main__":
    main()
