#!/bin/bash
# GPU-box session: quality pipeline tests, qbench (default + occupancy variants), ncu launch lists
# and full captures of fmq_main_kernel -> gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_quality_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_quality.log
tail -4 gpurun_out/pytest_quality.log
for c in quality_cfg3 quality_cfg1 quality_cfg2; do
  timeout 300 python tools/qbench.py --config $c > gpurun_out/qbench_$c.json 2> gpurun_out/qbench_$c.err || tail -5 gpurun_out/qbench_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/qbench_$c.json"))
print("$c", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"], "share", d["roofline"]["kernel_share_of_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "err", d["cpu_baseline"] and d["cpu_baseline"]["max_abs_err_vs_gpu"])
PY
done
for v in q3 q4 q6; do
  [ -f variants/$v/libfmdemod.so ] || continue
  for c in quality_cfg3 quality_cfg2; do
    FMD_LIB=$PWD/variants/$v/libfmdemod.so timeout 300 python tools/qbench.py --config $c --no-cpu > gpurun_out/qbench_${c}_$v.json 2> gpurun_out/qbench_${c}_$v.err
    python - <<PY
import json
d=json.load(open("gpurun_out/qbench_${c}_$v.json"))
print("$v $c", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"])
PY
  done
done
for c in quality_cfg3 quality_cfg2; do
  CMD="python tools/qbench.py --config $c --steps 2 --warmup 3 --no-cpu"
  $CMD > gpurun_out/qplain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/qlaunches_$c.csv $CMD > gpurun_out/qncu_list_$c.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:fmq_main -s 3 -c 1 -f -o gpurun_out/qprof_$c $CMD > gpurun_out/qncu_full_$c.log 2>&1
  tail -2 gpurun_out/qncu_full_$c.log
done
