#!/bin/bash
# 8-GPU box, final tree: the bench line at N = 8 (all legs) and the headline + e2e + same-run copy ceiling at N = 4, 2, 1
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
run() { # N extra-flags...
  local n=$1; shift
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $O/r02_scale_n$n.json 2> $O/r02_scale_n$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > $O/r02_scale_n$n.json 2> $O/r02_scale_n$n.err; fi
  echo "N=$n rc=$?"
}
run 8
run 4 --no-sublegs --no-sustained --no-cpu
run 2 --no-sublegs --no-sustained --no-cpu
run 1 --no-sublegs --no-sustained --no-cpu
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r02_scale_n{n}.json").read())
    except Exception as e:
        print(n, "no line:", e); continue
    e = d["e2e"]
    print(f"N={n}: device {d['value']/1e6:.3f} TS/s frac {d['roofline']['frac']}  e2e {e['value']/1e3:.2f} GS/s = {e['gbs_per_direction']} GB/s/dir, "
          f"ceiling {e['ceiling']['gbs_per_direction']} -> {e['frac_of_ceiling']}  parity {d['checks']['parity_max_abs']:.2e}")
PY
