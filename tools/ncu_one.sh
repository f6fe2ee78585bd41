#!/bin/bash
# usage: tools/ncu_one.sh <cfg> <outname>  -- plain run then one ncu --set full capture of the fast kernel
cfg=$1; out=$2
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --config $cfg"
$CMD > gpurun_out/plain_$out.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fmd_fast -s 30 -c 1 -f -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
tail -2 gpurun_out/ncu_$out.log
