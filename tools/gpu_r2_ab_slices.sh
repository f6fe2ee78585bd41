#!/bin/bash
# N GPUs (default 2): e2e with the round-2 default slices (64 MiB, graded) against the earlier form (32 MiB, equal), same box
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
O=gpurun_out; mkdir -p $O
run() { # tag env...
  local tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-sublegs --no-sustained --no-cpu > $O/r02_ab_slices_${tag}_n$N.json 2> $O/r02_ab_slices_${tag}_n$N.err
  echo "$tag rc=$?"
}
run new X=1
run old FMD_HOST_SLICE_MB=32 FMD_HOST_RAMP=0
run new2 X=1
python - <<PY
import json
for tag in ("new", "old", "new2"):
    try:
        d = json.loads(open(f"gpurun_out/r02_ab_slices_{tag}_n$N.json").read())
    except Exception as e:
        print(tag, "no line:", e); continue
    e = d["e2e"]
    print(f"{tag:5s} N=$N: device {d['value']/1e6:.3f} TS/s  e2e {e['value']/1e3:.2f} GS/s = {e['gbs_per_direction']} GB/s/dir, ceiling {e['ceiling']['gbs_per_direction']} -> {e['frac_of_ceiling']}  e2e==device {e['matches_device_path']}")
PY
