#!/usr/bin/env python
"""Per-kernel CUDA-event times of the bounded C4 / C5 / C3 workloads of bench.py (ctypes mirror)."""
import io
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import sequali_b200 as sq
    from sequali_b200 import _lib, synth
    ctx = _lib.Context.get()
    which = sys.argv[1:] or ["C4", "C5"]
    n = int(os.environ.get("N_READS", "20000"))
    for name in which:
        if name == "C4":
            data = synth.nanopore_fastq(n, mean_length=20_000, max_length=1_000_000, seed=4)
            run = lambda: bench.loop_single_end(sq, io.BytesIO(data), bench.NANOPORE_ADAPTERS, buffersize=64 << 20)
        elif name == "C5":
            data = synth.nanopore_ubam(n, mean_length=20_000, max_length=1_000_000, seed=5)
            run = lambda: bench.loop_single_end(sq, io.BytesIO(data), bench.NANOPORE_ADAPTERS, bam=True, buffersize=64 << 20)
        else:
            t1, t2 = synth.paired_fastq(400_000, seed=3)
            run = lambda: bench.loop_paired(sq, io.BytesIO(t1), io.BytesIO(t2), buffersize=64 << 20)
        run()
        t0 = time.perf_counter()
        bases = run()
        ctx.sync()
        dt = time.perf_counter() - t0
        ctx.profile(True)
        run()
        prof = ctx.profile_report()
        ctx.profile(False)
        ks = {k: v for k, v in prof.items() if not k.startswith("gap>")}
        gaps = sum(v[1] for k, v in prof.items() if k.startswith("gap>"))
        print(f"== {name}: {bases / dt / 1e9:.3f} Gbases/s, {dt * 1e3:.1f} ms wall; kernels {sum(v[1] for v in ks.values()):.1f} ms, "
              f"gaps {gaps:.1f} ms")
        for k, v in sorted(ks.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f"   {k:28s} {v[0]:5d} launches {v[1]:9.3f} ms")


if __name__ == "__main__":
    main()
