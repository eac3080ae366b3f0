#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --set full` capture as a small JSON that bench.py reads for its roofline block
(DRAM traffic and executed warp instructions per frame), so that the numbers on the bench line come from a committed
profile and not from literals.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/rNN_ncu_full_raw.csv
    python tools/ncu_kernels_json.py profiles/rNN_ncu_full_raw.csv profiles/rNN_kernels.json

Frames covered by the capture = sum over the k_describe_tile launches of their grid's y extent (one CTA row per frame).
"""
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "inst": 1.0, "": 1.0}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = {"time_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum", "dram_write_bytes": "dram__bytes_write.sum",
            "warp_inst": "smsp__inst_executed.sum"}
    kernels, frames = {}, 0
    body = rows[2:]
    # whole launch sequences only: a capture cut by -c may end with the first kernels of a sequence whose describe
    # launch (which tells how many frames it covered) is missing
    last = max((i for i, r in enumerate(body) if r[col["Kernel Name"]].startswith("k_describe")), default=len(body) - 1)
    for r in body[:last + 1]:
        name = r[col["Kernel Name"]].split("(")[0].strip()
        grid = [int(x) for x in re.findall(r"\d+", r[col["Grid Size"]])]
        k = kernels.setdefault(name, {"launches": 0, **{w: 0.0 for w in want}})
        k["launches"] += 1
        for w, c in want.items():
            i = col[c]
            k[w] += float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        if name.startswith("k_describe") and len(grid) >= 2:
            frames += grid[1]
    out = {"source": src, "frames": frames, "kernels": kernels}
    if frames:
        out["per_frame"] = {name: {w: k[w] / frames for w in want} for name, k in kernels.items()}
        out["per_frame_total"] = {w: sum(k[w] for k in kernels.values()) / frames for w in want}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out.get("per_frame_total")), frames)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
