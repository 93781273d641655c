#!/usr/bin/env python
"""Compact per-launch table of an `ncu --set full` report (for profiles/*.md):

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/prof.csv ; python tools/ncu_brief.py /tmp/prof.csv
"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("nunerf::", "")
        cells = []
        for key, _ in COLS:
            if key not in hdr:
                cells.append("-")
                continue
            i = hdr.index(key)
            v, u = r[i], units[i]
            try:
                f = float(v.replace(",", ""))
                if u == "ms":
                    f, u = f * 1e3, "us"
                v = f"{f:.1f}" if abs(f) < 1000 and not f.is_integer() else f"{f:.0f}"
            except ValueError:
                pass
            u = {"Mbyte": " MB", "Kbyte": " KB", "Gbyte": " GB", "byte": " B", "us": " us", "ms": " ms", "%": "",
                 "register/thread": ""}.get(u, "")
            cells.append(v + u)
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
