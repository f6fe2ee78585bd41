#!/bin/bash
# 1 GPU: what costs fmd_demod_host ~40 us of link time per slice?  An empty kernel in the copy-only form; the kernels on the
# (FMD_HOST_KSTREAM and FMD_COPY_NULLK were probe knobs of commit 8af8774; removed again after this measurement)
# copy-in / copy-out stream instead of a stream of their own.  (32 MiB equal slices make the per-slice cost most visible.)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
{
for geo in "FMD_HOST_SLICE_MB=32 FMD_HOST_RAMP=0" "X=1"; do
  echo "== $geo"
  env $geo timeout 120 python tools/e2e_probe.py
  echo -n "copy-only with an empty kernel:   "; env $geo FMD_COPY_NULLK=1 timeout 120 python tools/e2e_probe.py
  echo -n "kernels on the copy-in stream:    "; env $geo FMD_HOST_KSTREAM=in FMD_COPY_NULLK=1 timeout 120 python tools/e2e_probe.py
  echo -n "kernels on the copy-out stream:   "; env $geo FMD_HOST_KSTREAM=out FMD_COPY_NULLK=1 timeout 120 python tools/e2e_probe.py
done
} > $O/r02_e2e_probe2.txt 2>&1
cat $O/r02_e2e_probe2.txt
