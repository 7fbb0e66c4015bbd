#!/usr/bin/env python3
"""DRAM traffic per launch of the bench's kernels, from ncu --set full reports of the bench workload -> profiles/rNN_traffic.json.
usage: tools/ncu_traffic.py out.json bytes_per_gpu commit report.ncu-rep [report2.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys
out, nbytes, commit, reps = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4:]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
res = {"bytes_per_gpu": nbytes, "commit": commit, "how": "ncu --set full --clock-control none on bench.py's workload; traffic = dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", os.path.abspath(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
        if "nan" in d["dram__bytes_read.sum"].lower() or "nan" in d["dram__bytes_write.sum"].lower():
            continue                      # a pass that lost its DRAM counters: capture that kernel again
        rd = float(d["dram__bytes_read.sum"]) * SCALE[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * SCALE[u["dram__bytes_write.sum"]]
        res[name] = {"traffic": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                     "duration_ms": float(d["gpu__time_duration.sum"]) * TIME.get(u["gpu__time_duration.sum"], 1.0), "report": os.path.basename(rep)}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
