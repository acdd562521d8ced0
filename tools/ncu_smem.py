#!/usr/bin/env python
"""Shared-memory wavefronts per source line from an .ncu-rep.  usage: tools/ncu_smem.py rep [samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; L = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "": continue
    d = dict(zip(hdr, r))
    try: wf = int(d["L1 Wavefronts Shared"]); ideal = int(d["L1 Wavefronts Shared Ideal"]); inst = int(d["Instructions Executed"])
    except (ValueError, KeyError): continue
    if wf: L.append((wf, ideal, inst, cur, d["Line No"], r[1].strip()[:80]))
tot = sum(l[0] for l in L)
print("total shared wavefronts", tot, "(", tot * 128 / 1e9, "GB )", "" if not nsamp else f"= {tot*128/nsamp:.1f} B/sample")
for l in sorted(L, key=lambda l: -l[0])[:25]:
    print(f"{100*l[0]/tot:6.2f}%  wf={l[0]:>10} ideal={l[1]:>10} inst={l[2]:>9}  " + (f"{l[0]*128/nsamp:5.1f} B/smp  " if nsamp else "") + f"{l[3]}:{l[4]}  {l[5]}")
