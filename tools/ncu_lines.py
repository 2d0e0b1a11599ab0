#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts to source lines.

  tools/ncu_lines.py <report.ncu-rep> <mangled kernel name> [launch index in the report] [cubin]

ncu's `--page source --csv` lists SASS instructions with their executed counts
but no line numbers; nvdisasm -g lists the same instructions in the same order
with `//## File "...", line N` markers.  Join the two by position."""
import csv
import io
import re
import subprocess
import sys
import collections


def main(rep, func, skip="0", cubin="/tmp/scratch/libsqgpu.1.sm_100a.cubin", top=40):
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
    lines, cur, inside = [], None, False
    for ln in dis.splitlines():
        if ln.lstrip().startswith(".section"):
            inside = (".text." + func) in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-s", str(skip), "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = []
    for r in rows[2:]:  # some ncu versions print the page twice: stop at the second header
        if r and r[0] == "Kernel Name":
            break
        body.append(r)
    wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    if len(body) != len(lines):
        print(f"warning: {len(body)} ncu rows vs {len(lines)} disassembled instructions", file=sys.stderr)
    agg = collections.Counter()
    samp = collections.Counter()
    wav = collections.Counter()
    for r, loc in zip(body, lines):
        agg[loc] += float(r[ii] or 0)
        samp[loc] += float(r[si] or 0)
        if wi is not None:
            wav[loc] += float(r[wi] or 0)
    tot, stot, wtot = sum(agg.values()), sum(samp.values()), max(sum(wav.values()), 1)
    src = {}
    print(f"total warp instructions {tot:.0f}, samples {stot:.0f}")
    for loc, n in agg.most_common(int(top)):
        text = ""
        if loc:
            if loc[0] not in src:
                try:
                    src[loc[0]] = open(f"/root/repo/sequali_b200/csrc/{loc[0]}").read().splitlines()
                except OSError:
                    src[loc[0]] = []
            if loc[1] - 1 < len(src[loc[0]]):
                text = src[loc[0]][loc[1] - 1].strip()
        print(f"{n / tot * 100:5.1f}% inst {samp[loc] / max(stot, 1) * 100:5.1f}% samples {wav[loc] / wtot * 100:5.1f}% smem-wavefronts  {loc}  {text[:90]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
