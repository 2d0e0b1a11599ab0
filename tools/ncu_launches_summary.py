#!/usr/bin/env python
"""Per-kernel summary (launches, total time, share) of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`; .gz accepted)."""
import collections
import csv
import gzip
import re
import sys


def main(path, title=""):
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rt", errors="replace") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rd:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        tot[name] += v * scale
        cnt[name] += 1
    keep = {k: v for k, v in tot.items() if not k.startswith("k_synth")}
    total = sum(keep.values())
    print(f"# {title}\n")
    print(f"{sum(cnt[k] for k in keep)} launches, {total / 1e3:.1f} ms of kernel time "
          "(cold-cache, serialised: compare shares, not absolutes; synthetic generator kernels left out)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(keep.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {cnt[k]} | {v:.1f} | {v / total:.3f} |")


if __name__ == "__main__":
    main(*sys.argv[1:])
