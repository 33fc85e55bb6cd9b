"""Warp-instruction and stall-sample shares of an ncu report grouped by source-line ranges:
python profiles/source_groups.py <rep> <kernel> file:lo-hi=name ...   (lines not covered go to 'other')"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
groups = []
for g in sys.argv[3:]:
    rng, name = g.split("=")
    f, r = rng.split(":")
    lo, hi = r.split("-")
    groups.append((f, int(lo), int(hi), name))
sel = ["--launch-skip", kern[5:]] if kern.startswith("skip:") else ["--kernel-name", kern]   # "skip:N" = the (N+1)-th launch of the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel + ["--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, acc, tot_s, tot_i = "", {}, 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    if len(r) > 8 and r[0].isdigit():
        try:
            s, ie = int(r[6]), int(r[7])
        except ValueError:
            continue
        ln = int(r[0]); name = "other:" + cur
        for f, lo, hi, nm in groups:
            if f == cur and lo <= ln <= hi:
                name = nm; break
        a = acc.setdefault(name, [0, 0]); a[0] += s; a[1] += ie
        tot_s += s; tot_i += ie
print(f"total samples {tot_s}  warp instructions {tot_i}")
for name, (s, ie) in sorted(acc.items(), key=lambda t: -t[1][1]):
    print(f"{name:32s} inst {ie / tot_i * 100:5.1f}%  samples {s / tot_s * 100:5.1f}%")
