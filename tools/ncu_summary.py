#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ quote.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [--sass]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__shared_mem_per_block",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    for v in vals:
        d = dict(zip(hdr, v)); u = dict(zip(hdr, units))
        print("## kernel:", d.get("Kernel Name"))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d: print(f"| {k} | {d[k]} | {u[k]} |")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                try:
                    if float(d[k]) > 0.08: print(f"| {k} | {d[k]} | {u[k]} |")
                except ValueError: pass
    if "--sass" in sys.argv:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        # find header
        for i, r in enumerate(rows):
            if "Source" in r and "# Instructions Executed" in " ".join(r) or ("Source" in r and "Instructions Executed" in r): hdr = r; start = i + 1; break
        else:
            print("no source page"); return
        si = hdr.index("Source"); ei = hdr.index("Instructions Executed")
        from collections import Counter
        c = Counter(); tot = 0
        for r in rows[start:]:
            if len(r) <= max(si, ei): continue
            try: n = int(r[ei])
            except ValueError: continue
            op = r[si].split()[0] if r[si].split() else "?"
            if op.startswith("@"): op = r[si].split()[1]
            op = op.split(".")[0]
            c[op] += n; tot += n
        print("SASS opcode mix (warp instructions, % of", tot, "):", ", ".join(f"{k} {100*v/tot:.1f}" for k, v in c.most_common(22)))
main()
