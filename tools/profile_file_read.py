#!/usr/bin/env python
"""FastqParser of the extension on a regular file in the page cache: the parser alone (read + copy + scan) and with
the collectors, for several read steps."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import sequali_b200.ext as sqx
    from sequali_b200 import _lib
    from sequali_b200.device import DeviceFastq
    n = int(os.environ.get("N_READS", "12000000"))
    ctx = _lib.Context.get()
    data = DeviceFastq.synth_illumina(n, bench.READ_LENGTH, seed=2, chunk_reads=1 << 22)
    host, _ = data.to_host()
    where = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    fd, path = tempfile.mkstemp(prefix="sq_read_", suffix=".fastq", dir=where)
    with os.fdopen(fd, "wb") as f:
        f.write(memoryview(host))
    nbytes = os.path.getsize(path)
    data.free()
    try:
        for step in (16 << 20, 64 << 20, 256 << 20):
            for what in ("parser only", "parser + collectors"):
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    mods = bench.make_modules(sqx) if what != "parser only" else None
                    reads = 0
                    with open(path, "rb") as f:
                        for arr in sqx.FastqParser(f, step):
                            reads += len(arr)
                            if mods:
                                bench.feed(mods, arr)
                    if mods:
                        bench.read_results(mods)
                    sqx._qc._sync()
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                print(f"step {step >> 20:4d} MiB  {what:20s} {nbytes / best / 1e9:6.2f} GB/s  {reads * bench.READ_LENGTH / best / 1e9:6.2f} Gbases/s")
    finally:
        os.unlink(path)


if __name__ == "__main__":
    main()
