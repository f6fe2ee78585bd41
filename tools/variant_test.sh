#!/bin/bash
# parity (fast subset) + bench for each variant library
for d in default variants/*/; do
  if [ "$d" = default ]; then unset FMD_LIB; name=default; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
  r=$(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1)
  echo "$name parity: $r"
  for c in cfg2 cfg1; do
    python bench.py --steps 10 --warmup 3 --no-cpu --config $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name $c kernel_ms', r['kernel_ms_avg'], 'frac', r['frac'], 'value', d['value'])"
  done
done
