#!/bin/bash
# 8 GPUs, graded-slice tree: headline + e2e + same-run ceiling only
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-sublegs --no-sustained --no-cpu > $O/r02_n8_slices.json 2> $O/r02_n8_slices.err
echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_n8_slices.json").read())
e = d["e2e"]
print(f"N=8: device {d['value']/1e6:.3f} TS/s frac {d['roofline']['frac']}  e2e {e['value']/1e3:.2f} GS/s = {e['gbs_per_direction']} GB/s/dir, "
      f"ceiling {e['ceiling']['gbs_per_direction']} -> {e['frac_of_ceiling']}  parity {d['checks']['parity_max_abs']:.2e} e2e==device {e['matches_device_path']}")
PY
