"""profiles/ncu_traffic.json from `ncu --set full` captures, by script (no hand-entered numbers).

For every entry of MANIFEST: read the .ncu-rep with `ncu -i <rep> --page raw --csv`, pick the launch of the named kernel,
convert the unit row to bytes / microseconds, and record dram__bytes_read.sum + dram__bytes_write.sum per launch (the
`roofline.traffic` figure of bench.py), the launch duration and the tensor-pipe / DRAM utilisation ncu reports.  The
selected metric columns of each capture are also written to profiles/ncu_raw/<key>.csv so that the JSON can be re-derived
without the binary report.

    python tools/ncu_traffic.py [--reps gpurun_out] [--out profiles/ncu_traffic.json]

Captures are made on the GPU box AFTER the same command exited 0 without ncu, one kernel per capture:
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/<name> python tools/<bench>.py
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# key -> (report file, kernel-name regex, workload note with the ALGORITHMIC bytes of that launch)
MANIFEST = {
    "mlp_chain_kernel": ("prof_chain_train_r1.ncu-rep", r"mlp_chain_kernel",
                         "tools/bench_pred.py once: fused predictor forward chain 128>256>256>256>16 with per-layer stores and "
                         "ReLU masks, 383 000 rows; algorithmic = input 98.0 MB + 3 activations 588.3 MB + masks 36.8 MB + "
                         "head 24.5 MB = 747.6 MB (the tail of the last stores is still in L2 when the kernel ends)"),
    "mlp_chain_kernel_sdf_infer_r2": ("prof_r2_chain_ss.ncu-rep", r"mlp_chain_kernel",
                                      "round 2, SS kernel: fused SDF inference chain (PE in-kernel, 9 layers, sdf head only) "
                                      "on 303 104 points; algorithmic = 3.6 MB points + 1.2 MB sdf (weights 1.2 MB from L2)"),
    "mlp_chain_ts_kernel_sdf_infer_r2": ("prof_r2_chain_ts.ncu-rep", r"mlp_chain_ts_kernel|chain_ts",
                                         "round 2, TS kernel (A operand in TMEM, NUNERF_CHAIN_IMPL=ts): same workload"),
    "linear_tc_kernel": ("prof_linear_r1.ncu-rep", r"linear_tc_kernel",
                         "tools/bench_linear.py, one 383 475 x 256 x 256 bias+relu+mask_out layer; algorithmic 2*M*(K+N) + mask "
                         "= 404.9 MB (the tail of the output is still in L2 when the kernel ends)"),
    "dw_tc_kernel": ("prof_dw_r1.ncu-rep", r"dw_tc_kernel", "383 475 x 256 x 256 dW; algorithmic 2*M*(K+N) = 392.7 MB"),
}

COLUMNS = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__warps_active.avg.per_cycle_active")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
         "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def value(cell, unit):
    v = float(cell.replace(",", ""))
    return v * SCALE.get(unit, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ncu_traffic.json"))
    a = ap.parse_args()
    result = {}
    if os.path.exists(a.out):
        result = json.load(open(a.out))
    for key, (rep, pattern, note) in MANIFEST.items():
        path = os.path.join(a.reps, rep)
        if not os.path.exists(path):
            print(f"[ncu_traffic] {rep} not found: keeping the committed entry for {key}", file=sys.stderr)
            continue
        head, units, rows = raw_rows(path)
        col = {c: i for i, c in enumerate(head)}
        rows = [r for r in rows if re.search(pattern, r[col["Kernel Name"]])]
        if not rows:
            print(f"[ncu_traffic] no launch matching {pattern} in {rep}", file=sys.stderr)
            continue
        r = rows[0]
        get = lambda c: value(r[col[c]], units[col[c]]) if c in col and r[col[c]] not in ("", "n/a") else None
        rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        result[key] = {
            "dram_bytes_per_launch": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
            "duration_us_under_ncu": get("gpu__time_duration.sum"),
            "tensor_pipe_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": get("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers_per_thread": get("launch__registers_per_thread"),
            "kernel": r[col["Kernel Name"]], "report": "gpurun_out/" + rep, "launches_in_report": len(rows),
            "note": note, "generated_by": "tools/ncu_traffic.py",
        }
        keep = [c for c in COLUMNS if c in col]
        with open(os.path.join(ROOT, "profiles", "ncu_raw", key + ".csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(keep)
            w.writerow([units[col[c]] for c in keep])
            for rr in rows:
                w.writerow([rr[col[c]] for c in keep])
    json.dump(result, open(a.out, "w"), indent=1)
    print(json.dumps({k: v["dram_bytes_per_launch"] for k, v in result.items()}))


if __name__ == "__main__":
    main()
