#!/usr/bin/env python3
"""Executed warp instructions and stall samples of one kernel per CUDA SOURCE line, with the opcodes behind each line, from an
.ncu-rep captured with --import-source on.  Usage: tools/ncu_src_lines.py rep [top]"""
import csv
import subprocess
import sys
from collections import Counter, defaultdict


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    h = rows[hi]
    n, iex, ismp = len(h), h.index("Instructions Executed"), h.index("# Samples")
    cur, src = None, {}
    ex, sm, ops = Counter(), Counter(), defaultdict(Counter)
    for r in rows[hi + 1:]:
        if r and r[0].isdigit():
            cur = int(r[0])
            src[cur] = r[1]
            continue
        if len(r) != n or not r[2] or r[2] == "...":
            continue
        try:
            e = float(r[iex] or 0)
        except ValueError:
            continue
        ex[cur] += e
        sm[cur] += float(r[ismp] or 0)
        s = r[3].split()
        ops[cur][(s[1] if s[0].startswith("@") else s[0]).split(".")[0]] += e
    tot, tots = sum(ex.values()), max(sum(sm.values()), 1)
    print(f"total warp instructions {tot:.0f}, samples {tots:.0f}")
    for ln, e in ex.most_common(top):
        mix = " ".join(f"{k}:{100 * v / tot:.1f}" for k, v in ops[ln].most_common(4))
        print(f"{ln:5d} {100 * e / tot:5.1f} % inst {100 * sm[ln] / tots:5.1f} % smp | {src[ln].strip()[:80]} | {mix}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
