#!/usr/bin/env python3
"""Per source line: instructions, stall samples, shared-memory wavefronts (actual / ideal), normalised per unit.
usage: tools/ncu_bylines.py report.ncu-rep file.cu units [min_inst_per_unit] [sort: line|inst|smem|stall]"""
import csv, io, os, subprocess, sys
rep = os.path.abspath(sys.argv[1]); fname = sys.argv[2]; units = float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 500
order = sys.argv[5] if len(sys.argv) > 5 else "line"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = {}; cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():                 # the per-line aggregate row (SASS rows have an empty first column)
        def f(k):
            try: return float(r[hdr.index(k)] or 0)
            except Exception: return 0.0
        key = (cur, int(r[0]))
        old = lines.get(key, (0, 0, 0, 0, ""))
        lines[key] = (old[0] + f("Instructions Executed"), old[1] + f("Warp Stall Sampling (All Samples)"),
                      old[2] + f("L1 Wavefronts Shared"), old[3] + f("L1 Wavefronts Shared Ideal"), r[1].strip()[:90])
tot = sum(v[0] for v in lines.values()); tots = sum(v[1] for v in lines.values()) or 1; totw = sum(v[2] for v in lines.values())
print(f"total inst {tot:.4g} = {tot/units:.0f}/unit   shared wavefronts {totw:.4g} = {totw/units:.0f}/unit")
keyf = {"line": lambda kv: kv[0], "inst": lambda kv: -kv[1][0], "smem": lambda kv: -kv[1][2], "stall": lambda kv: -kv[1][1]}[order]
for k, (v, s, w, wi, t) in sorted(lines.items(), key=keyf):
    if (v / units >= thr or (order == "smem" and w / units >= thr)) and (k[0] == fname or fname == "all"):
        print(f"{k[0][:12]:12s}{k[1]:4d} {v/units:8.1f} inst {100*s/tots:4.1f}% stall  smem {w/units:7.1f} (ideal {wi/units:6.1f}) | {t}")
