#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, no GPU needed) into small text files for profiles/.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/summary.txt"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^launch__registers_per_thread$", r"^launch__grid_size$", r"^launch__block_size$",
    r"^launch__occupancy_limit", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__t_sector_hit_rate\.pct$", r"^lts__t_requests_srcunit_tex_op_read\.sum$",
    r"^lts__t_sectors_srcunit_tex_op_read\.sum$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_alu\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__average_warp.*_per_issue_stalled_.*\.ratio$", r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
    r"^sm__cycles_elapsed\.avg$", r"^sm__cycles_elapsed\.avg\.per_second$",
]


def main():
    pats = [re.compile(k) for k in KEYS]
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("== %s: unreadable" % rep); continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            print("== %s :: %s  grid=%s block=%s" % (rep.split("/")[-1], d.get("Kernel Name", "?")[:110], d.get("Grid Size", ""), d.get("Block Size", "")))
            for h, u, v in zip(hdr, units, vals):
                if any(p.search(h) for p in pats):
                    try:
                        if float(v) == 0 and "stalled" in h:
                            continue
                    except ValueError:
                        pass
                    print("   %-88s %18s %s" % (h, v, u))
            print()


if __name__ == "__main__":
    main()
