#!/usr/bin/env python
"""The BAM record chain: host walk (sq_bam_walk) against the device walk (sq_bam_walk_device, copy included), and
sq_batch_from_bam_bytes per kernel, on a short-read uBAM (many small records) and a nanopore uBAM (few long ones)."""
import ctypes as C
import os
import struct
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def short_read_ubam(n, length=150, seed=7):
    """n unaligned records of one size, assembled as a matrix (no Python loop per record)."""
    rng = np.random.default_rng(seed)
    name_w = 12
    tags = b"RGZrg1\0"
    body = 32 + name_w + (length + 1) // 2 + length + len(tags)
    rec = np.zeros((n, 4 + body), dtype=np.uint8)
    fixed = struct.pack("<IiiBBHHHIiii", body, -1, -1, name_w, 0, 4680, 0, 4, length, -1, -1, 0)
    rec[:, :36] = np.frombuffer(fixed, dtype=np.uint8)
    names = np.char.zfill(np.arange(n).astype("S11"), 11)
    rec[:, 36:36 + 11] = np.frombuffer(names.tobytes(), dtype=np.uint8).reshape(n, 11)
    p = 36 + name_w
    nib = rng.integers(0, 4, size=(n, 2 * ((length + 1) // 2)), dtype=np.uint8)
    nib = np.array([1, 2, 4, 8], dtype=np.uint8)[nib]
    rec[:, p:p + (length + 1) // 2] = nib[:, 0::2] << 4 | nib[:, 1::2]
    p += (length + 1) // 2
    rec[:, p:p + length] = rng.integers(2, 41, size=(n, length), dtype=np.uint8)
    p += length
    rec[:, p:] = np.frombuffer(tags, dtype=np.uint8)
    return rec.tobytes()


def main():
    from sequali_b200 import _lib, synth
    ctx = _lib.Context.get()
    lib = ctx.lib
    hdr = len(synth.bam_header())
    cases = {"short 2M x 150": short_read_ubam(int(os.environ.get("N_SHORT", "2000000"))),
             "nanopore 20k": synth.nanopore_ubam(int(os.environ.get("N_LONG", "20000")), 20_000, 1_000_000, seed=5)[hdr:]}
    for name, raw in cases.items():
        n = len(raw)
        from sequali_b200._qc import _PinnedBuffer
        pinned = _PinnedBuffer(ctx, n)
        C.memmove(pinned.ptr, raw, n)
        offs = np.zeros(n // 36 + 2, dtype=np.uint64)
        kept, skipped, used = C.c_uint64(), C.c_uint64(), C.c_uint64()
        res = {}
        for what in ("host", "device"):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                if what == "host":
                    rc = lib.sq_bam_walk(pinned.ptr, n, offs.ctypes.data, len(offs), C.byref(kept), C.byref(skipped), C.byref(used))
                else:
                    rc = lib.sq_bam_walk_device(ctx.h, pinned.ptr, n, 0, offs.ctypes.data, len(offs), C.byref(kept),
                                                C.byref(skipped), C.byref(used))
                best = min(best, time.perf_counter() - t0)
                assert rc == 0
            res[what] = (best, kept.value, used.value, int(offs[:kept.value].sum()))
        assert res["host"][1:] == res["device"][1:], res
        print(f"== {name}: {n / 1e6:.0f} MB, {res['host'][1]} records; host walk {res['host'][0] * 1e3:.1f} ms "
              f"({n / res['host'][0] / 1e9:.1f} GB/s), device walk incl. copy + offsets back {res['device'][0] * 1e3:.1f} ms "
              f"({n / res['device'][0] / 1e9:.1f} GB/s)")
        h, plen = C.c_void_p(), C.c_uint64()
        times = []
        for rep in range(5):
            if rep == 4:
                ctx.profile(True)
            t0 = time.perf_counter()
            rc = lib.sq_batch_from_bam_bytes(ctx.h, pinned.ptr, n, 0, C.byref(h), C.byref(kept), C.byref(skipped), C.byref(used),
                                             C.byref(plen))
            times.append(time.perf_counter() - t0)
            assert rc == 0
            lib.sq_batch_free(h)
        dt = min(times[:4])
        prof = ctx.profile_report()
        ctx.profile(False)
        ks = {k: v for k, v in prof.items() if not k.startswith("gap>")}
        print("   wall per call (last one profiled):", " ".join(f"{t * 1e3:.1f}" for t in times), "ms")
        print(f"   sq_batch_from_bam_bytes: {dt * 1e3:.1f} ms ({n / dt / 1e9:.1f} GB/s of BAM); kernels "
              f"{sum(v[1] for v in ks.values()):.2f} ms")
        for k, v in sorted(ks.items(), key=lambda kv: -kv[1][1])[:8]:
            print(f"      {k:24s} {v[0]:4d} launches {v[1]:8.3f} ms")


if __name__ == "__main__":
    main()
