#!/usr/bin/env python3
"""Instruction counts per source line in file order, normalised per block.
usage: tools/ncu_bylines.py report.ncu-rep file.cu nblocks [min_per_block]"""
import csv, subprocess, io, os, sys
rep = os.path.abspath(sys.argv[1]); fname = sys.argv[2]; nblk = float(sys.argv[3]); thr = float(sys.argv[4]) if len(sys.argv) > 4 else 500
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
lines={}; cur=None; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; hdr=None; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0].isdigit():
        def f(k):
            try: return float(r[hdr.index(k)] or 0)
            except Exception: return 0.0
        lines[(cur,int(r[0]))]=(f("Instructions Executed"), f("Warp Stall Sampling (All Samples)"), r[1].strip()[:100])
tot=sum(v[0] for v in lines.values()); tots=sum(v[1] for v in lines.values()) or 1
print(f"total inst {tot:.4g}  per block {tot/nblk:.0f}")
for k,(v,s,t) in sorted(lines.items()):
    if v/nblk >= thr and (k[0]==fname or fname=="all"):
        print(f"{k[0][:12]:12s}{k[1]:4d} {v/nblk:8.0f}/blk stall {100*s/tots:4.1f}% | {t}")
