#!/bin/bash
# one GPU-box session: parity tests, benches, ncu captures -> gpurun_out/
# usage (on the GPU box, from the repo root): tools/gpu_round.sh [ref] [ncu]
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for c in cfg2 cfg1 cfg3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --config $c $( [ $c != cfg2 ] && echo --no-cpu ) > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json"))
    print("$c", d["value"], "MS/s  roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms_avg"], "share", d["roofline"]["kernel_share_of_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "clk", d["clocks"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/bench_$c.err").read()[-2000:])
PY
done
if [[ " $* " == *" ref "* ]]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
fi
if [[ " $* " == *" ncu "* ]]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  for c in cfg2 cfg1; do
    CMDC="python bench.py --steps 2 --warmup 3 --no-cpu --config $c"
    $CMDC > gpurun_out/plain_$c.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:fmd_fast -s 3 -c 1 -f -o gpurun_out/prof_$c $CMDC > gpurun_out/ncu_full_$c.log 2>&1
    tail -2 gpurun_out/ncu_full_$c.log
  done
fi
