#!/usr/bin/env python3
"""Per-source-line summary of an ncu report: instructions executed and warp-stall samples.
usage: tools/ncu_lines.py report.ncu-rep [top_n] [--sass LINE]"""
import csv, subprocess, sys, io, os
rep = os.path.abspath(sys.argv[1]); top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
sass_line = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []; cur = None; hdr = None; last = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if not hdr: continue
    def f(k):
        i = hdr.index(k)
        try: return float(r[i] or 0)
        except (ValueError, IndexError): return 0.0
    if r[0].isdigit():
        last = (cur, int(r[0]))
        lines.append([cur, int(r[0]), r[1].strip()[:100], f("Instructions Executed"), f("Warp Stall Sampling (All Samples)")])
    elif sass_line is not None and last and last[1] == sass_line and r[0] == "":
        print(f"   {r[3].strip():60s} inst {f('Instructions Executed'):.3g} stall {f('Warp Stall Sampling (All Samples)'):.0f}")
tot_i = sum(l[3] for l in lines) or 1; tot_s = sum(l[4] for l in lines) or 1
print(f"total inst {tot_i:.4g}  total stall samples {tot_s:.0f}")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{l[0]}:{l[1]:4d} inst {100*l[3]/tot_i:5.1f}%  stall {100*l[4]/tot_s:5.1f}%  | {l[2]}")
