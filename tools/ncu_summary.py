#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [regex]"""
import csv, io, re, subprocess, sys

KEYS = [
    r"gpu__time_duration\.sum", r"dram__bytes_read\.sum$", r"dram__bytes_write\.sum$", r"gpu__dram_throughput\.avg\.pct",
    r"sm__throughput\.avg\.pct", r"sm__warps_active\.avg\.pct_of_peak_sustained_active", r"launch__registers_per_thread",
    r"launch__occupancy_limit", r"launch__grid_size", r"launch__block_size", r"smsp__inst_executed\.sum$",
    r"smsp__issue_active\.avg\.pct", r"sm__inst_executed_pipe_(fma|alu|lsu|fp64|xu|fmaheavy)\b.*sum$",
    r"smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__average_warp_latency_issue_stalled",
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"lts__t_sector_hit_rate\.pct", r"l1tex__t_sector_hit_rate\.pct",
    r"smsp__thread_inst_executed_per_inst_executed\.ratio", r"sm__cycles_elapsed\.max", r"lts__t_bytes\.sum$",
    r"smsp__cycles_active\.avg$", r"sm__pipe_.*cycles_active\.avg\.pct_of_peak_sustained_active",
    r"smsp__inst_executed_pipe_.*\.sum$",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:] if len(sys.argv) > 2 else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS + extra]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(p.search(h) for p in pats):
                print(f"  {h} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main()
