#!/usr/bin/env python3
"""Per-source-line share of executed warp instructions and stall samples of the first kernel in an .ncu-rep
(needs -lineinfo and --import-source on).  Usage: tools/ncu_lines.py report.ncu-rep [min_pct [kernel_name]]"""
import collections
import csv
import subprocess
import sys


def main(path, min_pct=0.5, kernel=None):
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kernel:
        cmd += ["--kernel-name", kernel, "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    for i, r in enumerate(rows):
        if "Source" in r and "Instructions Executed" in r:
            hdr, start = r, i + 1
            break
    li, ii, si, src = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    agg = collections.defaultdict(lambda: [0, 0, ""])
    tot = tots = 0
    for r in rows[start:]:
        if len(r) <= ii:
            continue
        try:
            ln, n, st = int(r[li]), int(r[ii]), int(r[si] or 0)
        except ValueError:
            continue
        agg[ln][0] += n
        agg[ln][1] += st
        agg[ln][2] = r[src][:100]
        tot += n
        tots += st
    print("total warp instructions", tot, "stall samples", tots)
    for ln, (n, st, s) in sorted(agg.items()):
        if n > tot * min_pct / 100 or st > tots * 2 * min_pct / 100:
            print(f"{ln:5d} {100 * n / tot:5.1f}% inst {100 * st / max(tots, 1):5.1f}% stall | {s}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5, sys.argv[3] if len(sys.argv) > 3 else None)
