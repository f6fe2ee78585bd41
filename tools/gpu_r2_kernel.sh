#!/bin/bash
# Round-2 GPU-box session (1 GPU): parity of the current build, A/B of the streaming-kernel variants
# under variants/, one ncu --set full capture of the cfg2 kernel.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
{
  for rep in 1 2; do
  for d in default variants/*/; do
    if [ "$d" = default ]; then unset FMD_LIB; name=default; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
    for c in cfg2 cfg1; do echo -n "$name: "; python tools/kbench.py $c 2048 40 2>&1 | tail -1; done
  done; done
  unset FMD_LIB
} > $O/r02_kernel_ab.txt 2>&1
python tools/kbench.py cfg2 2048 5 > $O/plain_kbench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fmd_fast_kernel -s 3 -c 1 -f -o $O/r02_fast_cfg2 \
    python tools/kbench.py cfg2 2048 5 > $O/ncu_kbench.log 2>&1
ncu -i $O/r02_fast_cfg2.ncu-rep --page raw --csv > $O/r02_fast_cfg2_raw.csv 2>/dev/null
ncu -i $O/r02_fast_cfg2.ncu-rep --page source --csv > $O/r02_fast_cfg2_source.csv 2>/dev/null
tail -3 $O/r02_pytest_gpu.log; cat $O/r02_kernel_ab.txt
