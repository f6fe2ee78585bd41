#!/bin/bash
# quality pipeline: tests with the default build, then qbench for the default and for every
# library under variants/ (device-resident leg only) -> gpurun_out/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_quality_gpu.py -m gpu -q 2>&1 | tail -6
QB() {
  python - <<PY
import json
d=json.load(open("gpurun_out/$1.json"))
print("$1", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"])
PY
}
for c in quality_cfg3 quality_cfg1 quality_cfg2 quality_m1; do
  timeout 300 python tools/qbench.py --config $c --no-cpu --no-e2e > gpurun_out/qv_${c}_default.json 2> gpurun_out/qv_${c}_default.err || tail -3 gpurun_out/qv_${c}_default.err
  QB qv_${c}_default
done
for v in $(ls variants); do
  for c in quality_cfg3 quality_cfg1 quality_cfg2 quality_m1; do
    FMD_LIB=$PWD/variants/$v/libfmdemod.so timeout 300 python tools/qbench.py --config $c --no-cpu --no-e2e > gpurun_out/qv_${c}_$v.json 2> gpurun_out/qv_${c}_$v.err || tail -3 gpurun_out/qv_${c}_$v.err
    QB qv_${c}_$v
  done
done
