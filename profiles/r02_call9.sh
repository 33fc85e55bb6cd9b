#!/usr/bin/env bash
# round 2, call 9: remaining GPU tests (default + checked build), launch list of the gzip path
set -uo pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest9.log
SCS_LIB_PATH=$PWD/scssim_b200/variants/libscssim_b200_checked.so python -m pytest tests -m gpu -x -q -k "not cli and not simuvars" > gpurun_out/r02_pytest9_checked.log 2>&1; echo "checked pytest rc=$?"; tail -4 gpurun_out/r02_pytest9_checked.log
CMD="python bench.py --steps 1 --warmup 1 --scale 0.02 --no-cpu-baseline --no-extras --gz"
$CMD > gpurun_out/r02_gz_plain.json 2> gpurun_out/r02_gz_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_gz_launches.csv $CMD > gpurun_out/r02_gz_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02_gz_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
acc=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); u=r[ui]
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    elif u=='s': v*=1e6
    k=r[ki].split('(')[0]
    acc[k][0]+=1; acc[k][1]+=v
tot=sum(v[1] for v in acc.values())
for k,(n,t) in sorted(acc.items(), key=lambda x:-x[1][1])[:14]: print(f"{t/1e3:9.2f} ms {t/tot*100:5.1f}% n={n:5d} avg {t/n:8.1f} us  {k[:70]}")
PY
