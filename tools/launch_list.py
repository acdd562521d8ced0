#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): tools/launch_list.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg, order = {}, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = d["Kernel Name"][:72]
        t = float(d["Metric Value"]) / 1e3
        if k not in agg:
            agg[k] = [0, 0.0, d["Grid Size"], d["Block Size"]]
            order.append(k)
        agg[k][0] += 1
        agg[k][1] += t
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | grid | block | total us | share |\n|---|---|---|---|---|---|")
for k in order:
    v = agg[k]
    print(f"| `{k}` | {v[0]} | {v[2]} | {v[3]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
print(f"| total | | | | {tot:.1f} | |")
