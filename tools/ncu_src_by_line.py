#!/usr/bin/env python3
"""Executed warp instructions and stall samples per SOURCE line (and per opcode) of one kernel of an .ncu-rep captured with
--import-source on.  Usage: tools/ncu_src_by_line.py rep [top]"""
import csv
import subprocess
import sys
from collections import Counter


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        # fall back to the plain source page
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h = rows[hi[0]]
    ci = {n: i for i, n in enumerate(h)}
    body = [r for r in rows[hi[0] + 1:] if len(r) > 10]
    ex = lambda r: float(r[ci["Instructions Executed"]] or 0)
    sm = lambda r: float(r[ci["# Samples"]] or 0)
    tot, tots = sum(map(ex, body)), sum(map(sm, body))
    print(f"total warp instructions {tot:.0f}, samples {tots:.0f}")
    ops, ops_s = Counter(), Counter()
    for r in body:
        src = r[ci["Source"]].split()
        op = (src[1] if src and src[0].startswith("@") else (src[0] if src else "?")).split(".")[0]
        ops[op] += ex(r)
        ops_s[op] += sm(r)
    for op, v in ops.most_common(18):
        print(f"  {op:10s} {100 * v / tot:5.1f} % of instructions  {100 * ops_s[op] / max(tots, 1):5.1f} % of samples")
    return body, ci


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
