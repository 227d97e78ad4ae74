#!/usr/bin/env python
"""usage: scripts/ncu_summary.py report.ncu-rep [header line ...]  - text summary of an `ncu --set full` report
(the metrics DESIGN.md / profiles/README.md quote, per kernel launch), read here with `ncu -i`."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__waves_per_multiprocessor"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for line in sys.argv[2:]:
        print("# " + line)
    for r in rows[2:]:
        d = dict(zip(head, r)); u = dict(zip(head, units))
        print("\n== " + d["Kernel Name"].split("(")[0])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]} {u.get(k, '')}")
        st = sorted(((float(d[k] or 0), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in d
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)[:4]
        print("  top stalls (warps per issue): " + ", ".join(f"{n} {v:.1f}" for v, n in st))


if __name__ == "__main__":
    main()
