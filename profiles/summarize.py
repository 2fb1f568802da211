#!/usr/bin/env python
"""profiles/summarize.py TAG -- condense the ncu outputs of profiles/capture.sh (in gpurun_out/) into the tracked
files profiles/<TAG>_ncu_summary.json, profiles/<TAG>_launch_shares.txt, profiles/<TAG>_launches.csv and refresh
profiles/kernel_traffic.json (read by bench.py for roofline.traffic)."""
import csv, json, os, re, shutil, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "gpurun_out")
dst = os.path.join(root, "profiles")

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
]

def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    out = {}
    for n, u, v in zip(names, units, vals):
        if n in WANT or n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio"):
            try:
                out[n] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                out[n] = {"value": v, "unit": u}
    out["kernel"] = vals[names.index("Kernel Name")]
    return out

summary, traffic = {}, {}
for k in ("k_encode", "k_decode", "k_fwd_stage_tma"):
    p = os.path.join(src, f"{tag}_{k}_raw.csv")
    if not os.path.exists(p):
        continue
    m = read_raw(p)
    stalls = {n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): v["value"]
              for n, v in m.items() if n.startswith("smsp__average_warps_issue_stalled")}
    m = {n: v for n, v in m.items() if not n.startswith("smsp__average_warps_issue_stalled")}
    m["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    def byt(name):
        v = m[name]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[v["unit"]]
        return v["value"] * mult
    traffic[k] = int(byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"))
    m["dram_bytes_total"] = traffic[k]
    summary[k] = m
json.dump(summary, open(os.path.join(dst, f"{tag}_ncu_summary.json"), "w"), indent=1)

old = json.load(open(os.path.join(dst, "kernel_traffic.json")))
old.update(traffic)
old["source"] = f"profiles/{tag}_ncu_summary.json: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, config 2, batch 1024)"
json.dump(old, open(os.path.join(dst, "kernel_traffic.json"), "w"), indent=1)

lp = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(dst, f"{tag}_launches.csv"))
    rows = list(csv.reader(l for l in open(lp) if l.startswith('"')))
    names = rows[0]
    ki, vi, ui = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        mult = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[r[ui]]
        name = re.sub(r"\(.*", "", r[ki])
        tot[name] += float(r[vi].replace(",", "")) * mult
        cnt[name] += 1
    allms = sum(tot.values())
    with open(os.path.join(dst, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e` ({tag}); times are cold-cache and\n"
                f"# serialised: read the SHARES, not the absolutes.  total {allms:.1f} ms over {sum(cnt.values())} launches\n")
        for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"{n:40s} {cnt[n]:5d} launches {t:10.3f} ms {100 * t / allms:6.2f} %  ({t / cnt[n]:.3f} ms each)\n")
print(open(os.path.join(dst, f"{tag}_launch_shares.txt")).read())
print(json.dumps({k: {"ms": v["gpu__time_duration.sum"], "dram": v["dram_bytes_total"], "stalls": v["top_stalls_per_issue"],
                      "thr/inst": v.get("smsp__thread_inst_executed_per_inst_executed.ratio"), "issue": v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                      "regs": v.get("launch__registers_per_thread"), "occ": v.get("sm__warps_active.avg.pct_of_peak_sustained_active")} for k, v in summary.items()}, indent=1))
