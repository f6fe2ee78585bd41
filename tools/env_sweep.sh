#!/bin/bash
# usage: tools/env_sweep.sh VAR v1 v2 ... : runs the short device-resident bench per value
var=$1; shift
for v in "$@"; do
  for c in cfg2 cfg1; do
    env $var=$v python bench.py --steps 10 --warmup 3 --no-cpu --config $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$var=$v $c kernel_ms', r['kernel_ms_avg'], 'frac', r['frac'], 'value', d['value'])"
  done
done
