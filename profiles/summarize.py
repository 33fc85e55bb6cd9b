"""Summarise gpurun_out/launches.csv (ncu launch list) and gpurun_out/prof.ncu-rep (ncu --set full) into
small text files that can be committed under profiles/.  usage: python profiles/summarize.py <tag>"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]

with open(f"profiles/{tag}_launches.txt", "w") as out:
    lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e6 if row["Metric Unit"] == "ns" else v / 1e3 if row["Metric Unit"] in ("us", "usecond") else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare shares)\n")
    out.write("# command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline  (3 resident + 3 e2e steps)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{k[:60]:60s} n={v[0]:5d} total={v[1]:9.2f} ms share={v[1] / tot * 100:5.1f}% avg={v[1] / v[0]:8.3f} ms\n")
    out.write(f"total {tot:.2f} ms\n")

raw = subprocess.run(["ncu", "-i", "gpurun_out/prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen = set()
with open(f"profiles/{tag}_ncu_full.txt", "w") as out:
    out.write("# ncu --set full --clock-control none --import-source on, one launch per kernel shown\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        if name in seen:
            continue
        seen.add(name)
        out.write(f"== {name}  grid={r[idx['Grid Size']]} block={r[idx['Block Size']]}\n")
        for k in KEYS:
            if k in idx:
                out.write(f"   {k:90s} {r[idx[k]]:>18s} {units[idx[k]]}\n")
# DRAM traffic of one emit launch (bench.py reports it as roofline.traffic)
import json
for r in rows[2:]:
    if r[idx["Kernel Name"]].startswith("emit_kernel") or "emit_kernel" in r[idx["Kernel Name"]]:
        def val(k):
            v = float(r[idx[k]].replace(",", "")); u = units[idx[k]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        json.dump({"kernel": "emit_kernel", "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                   "source": f"profiles/{tag}_ncu_full.txt (ncu --set full, python bench.py --steps 2 --warmup 1 --no-cpu-baseline)"},
                  open("profiles/emit_traffic.json", "w"))
        break
print(open(f"profiles/{tag}_launches.txt").read())
