#!/bin/bash
# 1 GPU, tree with the 64 MiB graded slices in fmd_demod_host: GPU suite, then the default bench line
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -3 $O/r02_pytest_gpu.log
timeout 600 python bench.py > $O/r02_bench_n1_slices.json 2> $O/r02_bench_n1_slices.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_slices.json").read())
e = d["e2e"]
print(f"device {d['value']/1e6:.3f} TS/s frac {d['roofline']['frac']}  e2e {e['value']/1e3:.2f} GS/s = {e['gbs_per_direction']} GB/s/dir, "
      f"ceiling {e['ceiling']['gbs_per_direction']} -> {e['frac_of_ceiling']}  parity {d['checks']['parity_max_abs']:.2e} e2e==device {e['matches_device_path']}")
PY
