#!/bin/bash
# experiment: the 8-outputs-per-lane streaming kernel (tools/experiments/fmd_fast2.cuh, variants/fast2)
export FMD_LIB=$PWD/variants/fast2/libfmdemod.so
FMD_FAST2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
for cfg in cfg2 cfg1; do
  echo "== $cfg baseline"; python tools/kbench.py $cfg 2048 20
  for ips in 4 8 16 32 64; do echo "== $cfg fast2 ips=$ips"; FMD_FAST2=1 FMD_IPS=$ips python tools/kbench.py $cfg 2048 20; done
done
