#!/bin/bash
# copy the judged artefacts of the quality pipeline from gpurun_out/ (scratch) into profiles/ (tracked)
R=${1:-r01}
cd "$(dirname "$0")/.."
for c in quality_cfg1 quality_cfg2 quality_cfg3 quality_m1; do cp gpurun_out/qbench_$c.json profiles/${R}_qbench_$c.json; done
cp gpurun_out/quality_cli_stream.txt profiles/${R}_quality_cli_stream.txt 2>/dev/null
for c in quality_cfg3 quality_cfg2; do
  cp gpurun_out/qlaunches_$c.csv profiles/${R}_quality_launches_$c.csv
  for k in fmq_main; do
    [ -f gpurun_out/qprof_${c}_${k}_raw.csv ] || continue
    cp gpurun_out/qprof_${c}_${k}_raw.csv profiles/${R}_${k}_${c}_raw.csv
    cp gpurun_out/qprof_${c}_${k}_details.txt profiles/${R}_${k}_${c}_details.txt
  done
done
python - <<PY > profiles/${R}_quality_ncu_summary.txt
import csv, collections, re
print("quality pipeline (FMD_FLAG_QUALITY) on B200: ncu evidence, tools/gpu_round4.sh")
print("command: python tools/qbench.py --config <c> --steps 2 --warmup 3 --no-cpu --no-e2e   (256 Mi samples per step)")
print()
print("launch lists (ncu --metrics gpu__time_duration.sum --clock-control none): serialised, cold-cache; compare SHARES")
for c in ("quality_cfg3", "quality_cfg2"):
    rows = list(csv.reader(l for l in open(f"profiles/${R}_quality_launches_{c}.csv") if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        m = re.search(r"(fmq_\w+<[^>]*>)", r[ki])
        if m: agg[m.group(1)].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"  {c}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"    {k:32s} {len(v):3d} launches  avg {sum(v)/len(v)/1e3:9.1f} us  share {sum(v)/tot:.3f}")
print()
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
STALL = 'smsp__average_warps_issue_stalled_'
import os
for f in ("fmq_main_quality_cfg3", "fmq_main_quality_cfg2"):
    p = f"profiles/${R}_{f}_raw.csv"
    if not os.path.exists(p): continue
    rows = list(csv.reader(open(p)))
    hdr, units, r = rows[0], rows[1], rows[2]
    print("==", f, "(ncu --set full --clock-control none, one launch over 256 Mi samples)")
    for k in KEYS:
        if k in hdr: print(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    st = [(float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i]]
    print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:8]))
PY
cat profiles/${R}_quality_ncu_summary.txt
python - <<PY
import csv, json, os
p = "profiles/traffic.json"
t = json.load(open(p)) if os.path.exists(p) else {}
for c in ("quality_cfg3", "quality_cfg2"):
    f = f"profiles/${R}_fmq_main_{c}_raw.csv"
    if not os.path.exists(f): continue
    rows = list(csv.reader(open(f))); h, u, r = rows[0], rows[1], rows[2]
    val = lambda k: float(r[h.index(k)]) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u[h.index(k)]]
    t[c] = {"samples_per_launch": 268435456, "dram_bytes_per_launch": val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
            "source": f"{f} (fmq_main_kernel: dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)"}
json.dump(t, open(p, "w"), indent=1)
PY
