#!/usr/bin/env python
"""Per-source-line stall samples and instruction counts from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0] == "": continue   # sass rows
    d = dict(zip(hdr, r))
    try:
        samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except (ValueError, KeyError): continue
    st = {k: int(d[k]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and d.get(k, "0").isdigit()}
    lines.append((samples, inst, cur_file, d["Line No"], hdr.index("Source") and r[1], st))
tot_s = sum(l[0] for l in lines); tot_i = sum(l[1] for l in lines)
print(f"total samples {tot_s}, total warp instructions {tot_i}")
agg = {}
for l in lines:
    for k, v in l[5].items(): agg[k] = agg.get(k, 0) + v
print("stall totals:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print(f"{'samp%':>6} {'inst%':>6}  file:line  source  | top stalls")
for l in sorted(lines, key=lambda l: -l[0])[:topn]:
    top = sorted(l[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*l[0]/tot_s:6.2f} {100*l[1]/tot_i:6.2f}  {l[2]}:{l[3]}  {l[4].strip()[:90]}  | " + ", ".join(f"{k[6:]}={v}" for k, v in top if v))
