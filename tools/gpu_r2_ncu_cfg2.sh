#!/bin/bash
# 1 GPU: ncu --set full of the final cfg2 and cfg1 kernels
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for c in cfg2 cfg1; do
python tools/kbench.py $c 2048 5 > $O/plain_kbench_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fmd_fast_kernel -s 3 -c 1 -f -o $O/r02_fast_$c \
    python tools/kbench.py $c 2048 5 > $O/ncu_kbench_$c.log 2>&1
ncu -i $O/r02_fast_$c.ncu-rep --page raw --csv > $O/r02_fast_kernel_${c}_raw.csv 2>/dev/null
done
python tools/ncu_summary.py $O/r02_fast_cfg2.ncu-rep $O/r02_fast_cfg1.ncu-rep > $O/r02_fast_kernel_ncu_summary.txt 2>/dev/null
cat $O/r02_fast_kernel_ncu_summary.txt | grep -E "==|duration|dram__bytes|lts__t_sectors|issue_active|stalls"
