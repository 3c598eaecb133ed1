"""Print the handful of ncu raw-page metrics we track for each captured launch of a .ncu-rep."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_uniform.sum",
]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("==", path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("  ", d.get("Kernel Name", "")[:60], d.get("Grid Size"))
        for w in WANT:
            if w in d:
                print(f"      {w:90s} {d[w]:>16s} {units[hdr.index(w)]}")
