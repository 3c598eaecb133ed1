"""profiles/rNN_traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list of one step (scripts/profile_step.py):   python scripts/make_traffic_json.py <csv> <out.json> "<source note>"
Per kernel family: launches, DRAM bytes read / written per step, bytes per launch (bench.py's roofline.traffic)."""
import collections
import csv
import io
import json
import sys

txt = open(sys.argv[1]).read().splitlines()
i = [n for n, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fam = collections.defaultdict(lambda: {"launch_ids": set(), "read": 0.0, "write": 0.0})
for r in rows:
    name = r["Kernel Name"]
    key = ("conv" if any(k in name for k in ("conv_tc_kernel", "conv_tc2_kernel", "conv_tcr_kernel", "conv_wn_kernel")) else
           "tail" if "tail_kernel" in name else "other")
    kern = next((k for k in ("conv_tcr_kernel", "conv_tc2_kernel", "conv_tc_kernel", "conv_wn_kernel") if k in name), None)
    f = fam[key]
    f["launch_ids"].add(r["ID"])
    if kern:
        f.setdefault("by_kernel", collections.Counter())
    v = float(r["Metric Value"].replace(",", "")) * unit_scale.get(r["Metric Unit"], 1.0)
    if r["Metric Name"] == "dram__bytes_read.sum":
        f["read"] += v
        if kern:
            f["by_kernel"][kern] += 1
    elif r["Metric Name"] == "dram__bytes_write.sum":
        f["write"] += v
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}
for key, f in fam.items():
    n = len(f["launch_ids"])
    out["conv_tc" if key == "conv" else key] = {
        "launches_per_step": n, "dram_read_bytes_per_step": f["read"], "dram_write_bytes_per_step": f["write"],
        "dram_bytes_per_launch": (f["read"] + f["write"]) / max(1, n),
        **({"kernels": ", ".join(f"{k} ({v})" for k, v in f["by_kernel"].items())} if "by_kernel" in f else {}),
    }
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
