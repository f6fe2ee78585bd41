#!/bin/bash
# 1 GPU: -q1 timings at full size, and where its time goes (ncu launch list)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python tools/conditioner_timing.py > $O/r02_conditioner_timing.txt 2>&1
timeout 300 python tools/conditioner_timing.py -q1 > $O/plain_q1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_q1.csv \
    python tools/conditioner_timing.py -q1 > $O/ncu_q1.log 2>&1
cat $O/r02_conditioner_timing.txt
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_launches_q1.csv")))
hdr = None; tot = collections.Counter(); n = collections.Counter()
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        name = r[hdr.index("Kernel Name")][:70]; v = float(r[hdr.index("Metric Value")].replace(",", "")); u = r[hdr.index("Metric Unit")]
        v *= {"usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6}.get(u, 1.0)
        tot[name] += v; n[name] += 1
for k, v in tot.most_common(8): print(f"{v/n[k]/1e3:9.1f} us avg x {n[k]:3d}  {k}")
PY
