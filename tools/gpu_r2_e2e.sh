#!/bin/bash
# 1 GPU: fmd_demod_host with slot reuse ordered on the device (default) vs by the host (round 1), plus the GPU suite
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
for rep in 1 2; do for m in 0 1; do
  echo -n "FMD_HOST_ORDERED=$m: "
  FMD_HOST_ORDERED=$m python bench.py --steps 10 --warmup 3 --no-cpu --no-sublegs --no-sustained 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e', e['value'], 'MS/s', e['gbs_per_direction'], 'GB/s per direction; ceiling', e['ceiling']['gbs_per_direction'], 'frac', e['frac_of_ceiling'])"
done; done > $O/r02_e2e_ordering.txt 2>&1
tail -3 $O/r02_pytest_gpu.log; cat $O/r02_e2e_ordering.txt
