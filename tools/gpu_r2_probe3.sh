#!/bin/bash
# 1 GPU: fewer resident CTAs for the kernels of a host-pipeline slot (FMD_HOST_PAD_SMEM_KB), then the GPU suite with and without it
# (FMD_HOST_PAD_SMEM_KB was a probe knob of commit ea42c97; it changed nothing and was removed again)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
{
for kb in 0 40 70 150 0 70; do
  echo -n "FMD_HOST_PAD_SMEM_KB=$kb: "; FMD_HOST_PAD_SMEM_KB=$kb PROBE_ORDER=dcd PROBE_STEPS=4 timeout 120 python tools/e2e_probe.py
done
} > $O/r02_e2e_probe3.txt 2>&1
cat $O/r02_e2e_probe3.txt
FMD_HOST_PAD_SMEM_KB=70 timeout 600 python -m pytest tests -q -m gpu -x > $O/r02_pytest_gpu_pad.log 2>&1; echo "pytest (padded) rc=$?"; tail -2 $O/r02_pytest_gpu_pad.log
timeout 600 python -m pytest tests -q -m gpu -x > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu.log
