#!/usr/bin/env python
"""Reads dram__bytes_read.sum + dram__bytes_write.sum of every kernel in an ncu report (`ncu --set full`) and writes /
updates profiles/dram_traffic.json, the file bench.py's roofline.traffic is taken from (a measurement, with its source).

    python tools/ncu_traffic.py gpurun_out/r02_dec2048.ncu-rep --batch 100000 --kernel-kind k_dec_pair
"""
import argparse
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--batch", type=int, required=True)
ap.add_argument("--kernel-kind", required=True, help="bench.py kernel kind the record is filed under (e.g. k_dec_pair)")
ap.add_argument("--match", default=None, help="substring of the kernel name in the report (default: the kind)")
args = ap.parse_args()
out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
recs = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    if (args.match or args.kernel_kind) not in d.get("Kernel Name", ""):
        continue
    tot = sum(float(d[k]) * UNIT[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    recs.append({"kernel": args.kernel_kind, "kernel_name": d["Kernel Name"], "batch": args.batch,
                 "dram_bytes_per_launch": tot, "dram_bytes_read": float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]],
                 "dram_bytes_write": float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]],
                 "gpu_time_ms": float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u["gpu__time_duration.sum"], 1.0),
                 "source": "profiles/" + os.path.basename(args.report).replace(".ncu-rep", "_raw.csv")})
path = os.path.join(ROOT, "profiles", "dram_traffic.json")
old = json.load(open(path)) if os.path.exists(path) else []
old = [r for r in old if not (r["kernel"] == args.kernel_kind and r["batch"] == args.batch)]
json.dump(old + recs[:1], open(path, "w"), indent=1)
with open(os.path.join(ROOT, "profiles", os.path.basename(args.report).replace(".ncu-rep", "_raw.csv")), "w") as f:
    f.write(out)
print(json.dumps(recs[:1], indent=1))
