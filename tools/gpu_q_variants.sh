#!/bin/bash
# quality pipeline: qbench (device-resident leg only) for the default build and for every library
# under variants/ -> gpurun_out/      usage: tools/gpu_q_variants.sh [config ...]
mkdir -p gpurun_out
CFGS=${@:-quality_cfg2}
QB() {
  python - <<PY
import json
d=json.load(open("gpurun_out/$1.json"))
print("$1", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"])
PY
}
for c in $CFGS; do
  timeout 300 python tools/qbench.py --config $c --no-cpu --no-e2e > gpurun_out/qv_${c}_default.json 2> gpurun_out/qv_${c}_default.err || tail -3 gpurun_out/qv_${c}_default.err
  QB qv_${c}_default
done
for v in $(ls variants); do
  for c in $CFGS; do
    FMD_LIB=$PWD/variants/$v/libfmdemod.so timeout 300 python tools/qbench.py --config $c --no-cpu --no-e2e > gpurun_out/qv_${c}_$v.json 2> gpurun_out/qv_${c}_$v.err || tail -3 gpurun_out/qv_${c}_$v.err
    QB qv_${c}_$v
  done
done
