#!/bin/bash
# 1 GPU: slice geometry of fmd_demod_host (size, graded start/end, slots) against the copy-only form of the same calls
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
{
timeout 120 python tools/e2e_probe.py
PROBE_ORDER=cdcd timeout 120 python tools/e2e_probe.py
FMD_HOST_RAMP=1 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=64 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=64 FMD_HOST_RAMP=1 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=128 FMD_HOST_RAMP=1 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=128 FMD_HOST_RAMP=1 FMD_HOST_SLOTS=4 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=16 FMD_HOST_RAMP=1 FMD_HOST_SLOTS=4 timeout 120 python tools/e2e_probe.py
FMD_HOST_SLICE_MB=256 FMD_HOST_RAMP=1 timeout 120 python tools/e2e_probe.py
} > $O/r02_e2e_probe.txt 2>&1
cat $O/r02_e2e_probe.txt
