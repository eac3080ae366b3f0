#!/usr/bin/env python3
"""Opcode histogram / region breakdown of one kernel from an .ncu-rep source page.
Usage: tools/ncu_sass_hist.py report.ncu-rep kernel_name [launch_index]"""
import collections
import csv
import subprocess
import sys


def main(path, kernel, units=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    for i, r in enumerate(rows):
        if "Source" in r and "Instructions Executed" in r:
            hdr, start = r, i + 1
            break
    si, ii, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    tot = 0
    region = 0
    reg = collections.defaultdict(int)
    regt = collections.defaultdict(int)
    ops = collections.defaultdict(lambda: collections.defaultdict(int))
    allops = collections.defaultdict(int)
    for r in rows[start:]:
        if len(r) <= ti or r[0] == "Kernel Name":
            if r and r[0] == "Kernel Name" and tot:
                break  # first launch only
            continue
        try:
            n, t = int(r[ii]), int(r[ti])
        except ValueError:
            continue
        s = r[si].strip()
        toks = s.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        reg[region] += n
        regt[region] += t
        ops[region][op] += n
        allops[op] += n
        tot += n
        if "BAR.SYNC" in s:
            region += 1
    print("total warp instr", tot)
    print("all:", ", ".join(f"{o}:{c / 1e6:.1f}M" for o, c in sorted(allops.items(), key=lambda kv: -kv[1])[:20]))
    for k in sorted(reg):
        print(f"region {k}: {reg[k] / 1e6:.1f}M {100 * reg[k] / tot:.1f}%  active lanes {regt[k] / max(1, reg[k]):.1f}")
        print("    ", ", ".join(f"{o}:{c / 1e6:.1f}M" for o, c in sorted(ops[k].items(), key=lambda kv: -kv[1])[:12]))


if __name__ == "__main__":
    main(*sys.argv[1:3])
