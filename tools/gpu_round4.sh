#!/bin/bash
# full GPU-box session after the quality pipeline landed: smoke, whole GPU suite, reference-mode
# bench lines, quality bench lines (+ the 64-samples-per-thread variant), ncu launch lists and
# --set full captures of the quality kernels (turned into CSV/text on the box), -Q stream through
# the CLI -> gpurun_out/
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for c in cfg2 cfg1 cfg3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --config $c $( [ $c != cfg2 ] && echo --no-cpu ) > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$c.json"))
    print("$c", d["value"], "MS/s  roofline", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms_avg"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/bench_$c.err").read()[-1500:])
PY
done
QB() { # name, extra env
  python - <<PY
import json
d=json.load(open("gpurun_out/$1.json"))
print("$1", d["value"], "MS/s  ms/step", d["ms_per_step"], "main_ms", d["roofline"]["kernel_ms_avg"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "err", d["cpu_baseline"] and d["cpu_baseline"]["max_abs_err_vs_gpu"])
PY
}
for c in quality_cfg3 quality_cfg1 quality_cfg2 quality_m1; do
  timeout 300 python tools/qbench.py --config $c > gpurun_out/qbench_$c.json 2> gpurun_out/qbench_$c.err || tail -5 gpurun_out/qbench_$c.err
  QB qbench_$c
done
for c in quality_cfg3 quality_cfg2; do
  CMD="python tools/qbench.py --config $c --steps 2 --warmup 3 --no-cpu --no-e2e"
  $CMD > gpurun_out/qplain_$c.log 2>&1 || { tail -5 gpurun_out/qplain_$c.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/qlaunches_$c.csv $CMD > gpurun_out/qncu_list_$c.log 2>&1
  for k in fmq_main; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/qprof_${c}_$k $CMD > gpurun_out/qncu_full_${c}_$k.log 2>&1
    ncu -i /tmp/qprof_${c}_$k.ncu-rep --page raw --csv > gpurun_out/qprof_${c}_${k}_raw.csv 2>/dev/null
    ncu -i /tmp/qprof_${c}_$k.ncu-rep --page details > gpurun_out/qprof_${c}_${k}_details.txt 2>/dev/null
  done
done
# one long -Q stream through the CLI: 4 GiB (2 Gi samples) through a stdin pipe and from a file
python - <<PY
import sys; sys.path.insert(0,'.')
from demodulator_b200 import synth
base = synth.config_signal("cfg3", 64)
with open("/dev/shm/fmd_stream.u8","wb") as f:
    for _ in range(64): f.write(base.tobytes())
PY
F=/dev/shm/fmd_stream.u8
{
echo "quality stream through the CLI: -Q -r2400000 -S48000 -l15000 (decimation 50), 1 GiB file = 512 Mi samples"
echo "--- file -> /dev/null"; FMD_STATS=1 demodulator_b200/bin/demodulator -i $F -o /dev/null -Q -r2400000 -S48000 -l15000
echo "--- stdin pipe (cat x4 = 2 Gi samples) -> /dev/null"; cat $F $F $F $F | FMD_STATS=1 demodulator_b200/bin/demodulator -i - -o /dev/null -Q -r2400000 -S48000 -l15000
echo "--- same, -Q -r192000 -S96000 -L12500 -l6500 (decimation 2, both cascades)"; cat $F $F $F $F | FMD_STATS=1 demodulator_b200/bin/demodulator -i - -o /dev/null -Q -r192000 -S96000 -L12500 -l6500
} > gpurun_out/quality_cli_stream.txt 2>&1
rm -f $F
cat gpurun_out/quality_cli_stream.txt
