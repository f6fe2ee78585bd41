#!/bin/bash
# 1-GPU check: whole GPU suite, conditioner timings, A/B of the early-LDS variant (burst + sustained)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
timeout 300 python tools/conditioner_timing.py > $O/r02_conditioner_timing.txt 2>&1
FMD_LIB=$PWD/variants/early/libfmdemod.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r02_pytest_early.log 2>&1; echo "pytest(early) rc=$?" >> $O/r02_pytest_early.log
{
  for rep in 1 2; do
  for d in default variants/early/; do
    if [ "$d" = default ]; then unset FMD_LIB; name=default; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
    for c in cfg2; do
      echo -n "$name burst: "; timeout 120 python tools/kbench.py $c 2048 40 2>&1 | tail -1
      echo -n "$name sustained 4400 launches: "; timeout 120 python tools/kbench.py $c 2048 4400 2>&1 | tail -1
    done
  done; done
  unset FMD_LIB
} > $O/r02_kernel_ab3.txt 2>&1
tail -6 $O/r02_pytest_gpu.log; cat $O/r02_conditioner_timing.txt; tail -2 $O/r02_pytest_early.log; cat $O/r02_kernel_ab3.txt
