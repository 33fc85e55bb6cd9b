"""Round-2 summaries: python profiles/summarize_r02.py <tag> <launches.csv> <report.ncu-rep>
-> profiles/<tag>_launches.txt, profiles/<tag>_ncu_full.txt, profiles/emit_traffic.json (DRAM bytes of one emit launch)"""
import collections, csv, json, re, subprocess, sys
tag, lcsv, rep = sys.argv[1:4]
CMD = "python bench.py --steps 2 --warmup 3 --scale 0.05 --no-extras --no-cpu-baseline"
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader([l for l in open(lcsv) if not l.startswith("==")]):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("scs::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    agg[name][0] += 1; agg[name][1] += v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
with open(f"profiles/{tag}_launches.txt", "w") as out:
    out.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none, launches 300..700 of `{CMD}`)\n# per-launch times are cold-cache and serialised: compare shares\n# total {tot:.2f} ms over {n} launches\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{v[1]:9.3f} ms {v[1] / tot * 100:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0] * 1e3:9.1f} us  {k[:70]}\n")
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
with open(f"profiles/{tag}_ncu_full.txt", "w") as out:
    out.write(f"# ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 30 -c 2, `{CMD}`; units as ncu prints them\n")
    for r in rows[2:]:
        out.write("launch: " + r[idx["Kernel Name"]][:60] + "\n")
        for k in KEYS:
            if k in idx: out.write(f"  {k} = {r[idx[k]]} {units[idx[k]]}\n")
r = rows[-1]
def val(k):
    return float(r[idx[k]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1)
json.dump({"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "source": f"profiles/{tag}_ncu_full.txt (ncu --set full, emit_kernel<0,0>, 191 k pairs per launch)",
           "read": val("dram__bytes_read.sum"), "write": val("dram__bytes_write.sum")}, open("profiles/emit_traffic.json", "w"))
print(open(f"profiles/{tag}_launches.txt").read()); print(open(f"profiles/{tag}_ncu_full.txt").read()[:1800]); print(open("profiles/emit_traffic.json").read())
