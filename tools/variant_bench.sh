#!/bin/bash
# run the short device-resident bench for every built variant: prints kernel ms and roofline frac
for d in default variants/*/; do
  if [ "$d" = default ]; then unset FMD_LIB; name=default; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
  for c in cfg2 cfg1; do
    python bench.py --steps 10 --warmup 3 --no-cpu --config $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$name $c kernel_ms', r['kernel_ms_avg'], 'frac', r['frac'], 'value', d['value'])"
  done
done
