#!/bin/bash
# 1 GPU: does the statically indexed exact path change the streaming kernel's speed? (same box, A/B)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
{
  for rep in 1 2 3; do
  for d in default variants/oldcone/; do
    if [ "$d" = default ]; then unset FMD_LIB; name=final; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
    for c in cfg2 cfg1; do echo -n "$name burst: "; timeout 120 python tools/kbench.py $c 2048 40 2>&1 | tail -1; done
  done; done
  unset FMD_LIB
} > $O/r02_kernel_ab4_exact_cone.txt 2>&1
cat $O/r02_kernel_ab4_exact_cone.txt
