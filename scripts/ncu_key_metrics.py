#!/usr/bin/env python
"""One table of the metrics that decide a kernel's bound, over every launch of the given .ncu-rep files (reads the reports
with `ncu -i ... --page raw --csv`; no GPU needed).

    python scripts/ncu_key_metrics.py gpurun_out/*.ncu-rep > profiles/r02_ncu_key_metrics.md
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STALL = "smsp__average_warps_issue_stalled_"


def num(d, k, default=float("nan")):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return default


def scale(unit):
    unit = unit.split("/")[0]
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


def main():
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        hbm = json.load(f)["hbm_gbs"]
    print("# Key ncu metrics of every full capture of round 2 (`ncu --set full --clock-control none`)\n")
    print(f"Read from the `.ncu-rep` files with `scripts/ncu_key_metrics.py`. tensor = `sm__pipe_tensor_cycles_active.avg.pct_of_peak_"
          f"sustained_active` and `..._elapsed` (the second counts idle SMs and idle time: the figure DESIGN.md quotes); issue = `sm__issue_active.avg.pct_of_peak_sustained_elapsed`; DRAM GB/s = (`dram__bytes_read.sum` + "
          f"`dram__bytes_write.sum`) / `gpu__time_duration.sum`, and its share of the measured copy peak ({hbm:.0f} GB/s, "
          "MEASURED_PEAKS.json); stalls = the three largest `smsp__average_warps_issue_stalled_*_per_issue_active` ratios. Times are "
          "of a profiled (serialised, cache-cold) launch.\n")
    print("| capture | kernel | grid | us | tensor % (active SM cycles) | tensor % (elapsed) | issue % | DRAM MB | DRAM GB/s | of HBM peak | L2 hit % | regs | smem KB | top stalls (per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        u = dict(zip(hdr, units))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            t_us = num(d, "gpu__time_duration.sum") * scale(u["gpu__time_duration.sum"])
            byt = num(d, "dram__bytes_read.sum") * scale(u["dram__bytes_read.sum"]) + \
                num(d, "dram__bytes_write.sum") * scale(u["dram__bytes_write.sum"])
            gbs = byt / t_us / 1e3
            stalls = sorted(((num(d, k, 0.0), k[len(STALL):-len("_per_issue_active.ratio")]) for k in hdr
                             if k.startswith(STALL) and k.endswith("_per_issue_active.ratio")), reverse=True)[:3]
            name = d["Kernel Name"].replace("void ", "").replace("unnamed>::", "").split("(")[0]
            print(f"| {os.path.basename(path).replace('.ncu-rep', '')} | `{name}` | {d['Grid Size']} | {t_us:.1f} | "
                  f"{num(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                  f"{num(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                  f"{num(d, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {byt / 1e6:.0f} | {gbs:.0f} | {gbs / hbm:.2f} | "
                  f"{num(d, 'lts__t_sector_hit_rate.pct'):.0f} | {num(d, 'launch__registers_per_thread'):.0f} | "
                  f"{num(d, 'launch__shared_mem_per_block_dynamic') * scale(u.get('launch__shared_mem_per_block_dynamic', 'byte')) / 1e3:.0f} | "
                  + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")


if __name__ == "__main__":
    main()
