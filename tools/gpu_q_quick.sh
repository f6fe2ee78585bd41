#!/bin/bash
# quick GPU check of the quality pipeline: its tests, then the qbench lines -> gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_quality_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_quality.log
tail -6 gpurun_out/pytest_quality.log
for c in quality_cfg2 quality_cfg3 quality_cfg1 quality_m1; do
  timeout 300 python tools/qbench.py --config $c $QB_ARGS > gpurun_out/qbench_$c.json 2> gpurun_out/qbench_$c.err || tail -5 gpurun_out/qbench_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/qbench_$c.json"))
print("$c", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"], "share", d["roofline"]["kernel_share_of_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "err", d["cpu_baseline"] and d["cpu_baseline"]["max_abs_err_vs_gpu"])
PY
done
