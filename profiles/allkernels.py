#!/usr/bin/env python
"""profiles/allkernels.py RAW.csv -- condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,...
--csv --page raw` launch list into one line per kernel: launches, total time, share of the step, DRAM bytes, GB/s vs the measured peak."""
import csv, json, os, sys
from collections import OrderedDict
path = sys.argv[1]
peak = 6582.5
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd); units = next(rd)
col = {n: i for i, n in enumerate(hdr)}
def num(r, name):
    i = col.get(name)
    if i is None: return 0.0
    try: v = float(r[i].replace(",", ""))
    except ValueError: return 0.0
    u = units[i].lower()
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    return v * scale
agg = OrderedDict()
for r in rd:
    if len(r) < len(hdr): continue
    k = r[col["Kernel Name"]].split("(")[0]
    a = agg.setdefault(k, dict(n=0, ms=0.0, rd=0.0, wr=0.0, inst=0.0, regs=0, grid=r[col["Grid Size"]], block=r[col["Block Size"]]))
    a["n"] += 1; a["ms"] += num(r, "gpu__time_duration.sum"); a["rd"] += num(r, "dram__bytes_read.sum"); a["wr"] += num(r, "dram__bytes_write.sum")
    a["inst"] += num(r, "smsp__inst_executed.sum"); a["regs"] = int(num(r, "launch__registers_per_thread"))
tot = sum(a["ms"] for a in agg.values())
print("%-34s %4s %9s %6s %10s %10s %8s %6s %5s %10s" % ("kernel", "n", "ms", "share", "dram_rd_MB", "dram_wr_MB", "GB/s", "frac", "regs", "warp_inst"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = (a["rd"] + a["wr"]) / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0
    print("%-34s %4d %9.3f %5.1f%% %10.1f %10.1f %8.1f %6.3f %5d %10.3g" % (k[:34], a["n"], a["ms"], 100 * a["ms"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs, gbs / peak, a["regs"], a["inst"]))
print("total %.3f ms (serialised, cold-cache ncu replay times; shares are what counts)" % tot)
