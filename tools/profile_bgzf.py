#!/usr/bin/env python
"""One pass of the BGZF reader over ~N MB of synthetic FASTQ text (for ncu on k_bgzf_inflate)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from sequali_b200 import _lib, synth
    from sequali_b200.device import HostFastq
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    text = synth.illumina_fastq(n_reads, length=150, seed=9, n_tiles=50)
    comp = synth.bgzf_compress(text, level=6)
    ctx = _lib.Context.get()
    host = HostFastq.from_bytes(comp)
    for rep in range(2):
        t0 = time.perf_counter()
        n = sum(len(a) for a in host.record_arrays_bgzf(256 << 20))
        ctx.sync()
        dt = time.perf_counter() - t0
        print(f"pass {rep}: {n} reads, {len(text) / dt / 1e9:.2f} GB/s of text, ratio {len(text) / len(comp):.2f}")
    ctx.profile(True)
    n = sum(len(a) for a in host.record_arrays_bgzf(256 << 20))
    prof = ctx.profile_report()
    ctx.profile(False)
    for k, v in prof.items():
        if "inflate" in k and not k.startswith("gap"):
            print(f"{k}: {v[0]} launches, {v[1]:.3f} ms -> {len(text) / v[1] / 1e6:.1f} GB/s of text")


if __name__ == "__main__":
    main()
