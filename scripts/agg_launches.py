"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel + grid."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read().splitlines()
i = [n for n, l in enumerate(txt) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(txt[i:]))))
per_launch = len(sys.argv) > 2
agg = collections.defaultdict(lambda: [0, 0])
for n, r in enumerate(rows):
    k = r['Kernel Name'].replace('void unnamed>::', '').replace('unnamed>::', '')[:44] + ' grid ' + r['Grid Size']
    if per_launch:
        print(n, k, int(r['Metric Value']) / 1e3)
    agg[k][0] += int(r['Metric Value'])
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
print("| kernel | launches | sum us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"| `{k}` | {v[1]} | {v[0] / 1e3:.1f} | {100 * v[0] / tot:.1f}% |")
print(f"total {tot / 1e3:.0f} us over {len(rows)} launches")
