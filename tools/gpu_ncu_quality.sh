#!/bin/bash
# ncu evidence for the quality pipeline: launch list + one --set full capture of fmq_main_kernel and of
# fmq_in_aggregate_kernel per config; the reports are turned into CSV/text here (gpurun_out/ is
# capped at 64 MiB) -> gpurun_out/
mkdir -p gpurun_out
for c in quality_cfg3 quality_cfg2; do
  CMD="python tools/qbench.py --config $c --steps 2 --warmup 3 --no-cpu"
  $CMD > gpurun_out/qplain_$c.log 2>&1 || { tail -5 gpurun_out/qplain_$c.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/qlaunches_$c.csv $CMD > gpurun_out/qncu_list_$c.log 2>&1
  for k in fmq_main fmq_in_aggregate; do
    [ $k = fmq_in_aggregate ] && [ $c = quality_cfg3 ] && continue
    ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/qprof_${c}_$k $CMD > gpurun_out/qncu_full_${c}_$k.log 2>&1
    ncu -i /tmp/qprof_${c}_$k.ncu-rep --page raw --csv > gpurun_out/qprof_${c}_${k}_raw.csv 2>/dev/null
    ncu -i /tmp/qprof_${c}_$k.ncu-rep --page details > gpurun_out/qprof_${c}_${k}_details.txt 2>/dev/null
    ncu -i /tmp/qprof_${c}_$k.ncu-rep --page source --csv > gpurun_out/qprof_${c}_${k}_source.csv 2>/dev/null
    tail -1 gpurun_out/qncu_full_${c}_$k.log
  done
done
ls -la gpurun_out | tail -20
