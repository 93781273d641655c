#!/usr/bin/env python
"""Pull the roofline-relevant counters of every profiled launch out of an `ncu --set full` report.

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/prof.csv ; python tools/ncu_metrics.py /tmp/prof.csv
"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if any(h == w or h.startswith(w) for w in WANT) or "pipe_tensor" in h]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"## {r[ki][:100]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for i in cols:
            print(f"- {hdr[i]} = {r[i]} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
