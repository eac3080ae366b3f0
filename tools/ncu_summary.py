#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into one line per profiled launch: the counters the
roofline discussion needs.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("smsp__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_lsb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "lsb/issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "bar/issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "ssb/issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "mpt/issue"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        parts = [r[ki].split("(")[0]]
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v):.4g}"
                except ValueError:
                    pass
                parts.append(f"{short}={v}{units[i] if units[i] not in ('%', '') and short in ('time','dram_rd','dram_wr') else ''}")
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
