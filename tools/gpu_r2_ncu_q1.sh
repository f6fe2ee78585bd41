#!/bin/bash
# 1 GPU: ncu --set full of the fused -q1 kernel (one launch over 512 Mi samples)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python tools/conditioner_timing.py "-q1 -S96000 -L" > $O/plain_q1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fmd_q1_fused_kernel -s 2 -c 1 -f -o $O/r02_q1_fused \
    python tools/conditioner_timing.py "-q1 -S96000 -L" > $O/ncu_q1.log 2>&1
ncu -i $O/r02_q1_fused.ncu-rep --page raw --csv > $O/r02_q1_fused_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02_q1_fused.ncu-rep > $O/r02_q1_fused_ncu_summary.txt 2>/dev/null
cat $O/plain_q1.log; cat $O/r02_q1_fused_ncu_summary.txt
