"""Per-CUDA-source-line stall samples from an ncu report: python profiles/source_hot.py <rep> <kernel> [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, out, tot, totinst = "", [], 0, 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) > 8 and r[0].isdigit():
        try:
            s, ie = int(r[6]), int(r[7])
        except ValueError:
            continue
        out.append((s, ie, cur_file, int(r[0]), r[1].strip()))
        tot += s; totinst += ie
print(f"total samples {tot}  warp instructions {totinst}")
for s, ie, f, ln, src in sorted(out, key=lambda t: -t[0])[:top]:
    print(f"{s:7d} {s / tot * 100:5.1f}%  inst {ie / totinst * 100:5.1f}%  {f}:{ln:<4d} {src[:110]}")
