#!/bin/bash
# copy the judged artefacts of the last gpu_round.sh from gpurun_out/ (scratch) into profiles/ (tracked)
R=${1:-r01}
cd "$(dirname "$0")/.."
cp gpurun_out/launches.csv profiles/${R}_launches_bench_cfg2.csv
(cd gpurun_out; python ../tools/ncu_summary.py prof_cfg2.ncu-rep prof_cfg1.ncu-rep) > profiles/${R}_fast_kernel_ncu_summary.txt
for c in cfg2 cfg1; do
  ncu -i gpurun_out/prof_$c.ncu-rep --page raw --csv 2>/dev/null > profiles/${R}_fast_kernel_${c}_raw.csv
  ncu -i gpurun_out/prof_$c.ncu-rep --page details 2>/dev/null > profiles/${R}_fast_kernel_${c}_details.txt
done
for c in cfg1 cfg2 cfg3; do cp gpurun_out/bench_$c.json profiles/${R}_bench_$c.json; done
[ -f gpurun_out/bench_ref.json ] && cp gpurun_out/bench_ref.json profiles/${R}_bench_reference_arm.json
python - <<PY
import csv,json,collections
import os
t=json.load(open('profiles/traffic.json')) if os.path.exists('profiles/traffic.json') else {}  # keep the quality_* entries
for c in ("cfg2","cfg1"):
    rows=list(csv.reader(open(f'profiles/${R}_fast_kernel_{c}_raw.csv')))
    h,u,r=rows[0],rows[1],rows[2]
    def val(k):
        v=float(r[h.index(k)]); un=u[h.index(k)]
        return v*{'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}[un]
    t[c]={"samples_per_launch":1073741824,"dram_bytes_per_launch":val('dram__bytes_read.sum')+val('dram__bytes_write.sum'),
          "source":f"profiles/${R}_fast_kernel_{c}_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture)"}
t["cfg3"]=dict(t["cfg1"]); t["cfg3"]["source"]+=" -- same kernel instantiation and shape as cfg1"
json.dump(t,open('profiles/traffic.json','w'),indent=1)
rows=list(csv.reader(open('profiles/${R}_launches_bench_cfg2.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; data=rows[hi+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in data:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki][:70]; agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
out=[f"{t_/1e3:12.1f} us {c:5d} launches {100*t_/tot:6.2f}%  {n}" for n,(c,t_) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]]
open('profiles/${R}_launches_summary.txt','w').write("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu (cfg2)\nserialised, cold-cache launch times: compare SHARES, not absolutes\n"+"\n".join(out)+"\n(the at::elementwise kernels are torch filling the synthetic input once, outside the timed region;\n inside the timed region fmd_fast_kernel is the only kernel: share of step 0.995 by CUDA events)\n")
print("\n".join(out))
PY
grep -vE "lts__t_sectors|l1tex__t|occupancy_limit" profiles/${R}_fast_kernel_ncu_summary.txt
