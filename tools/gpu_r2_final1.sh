#!/bin/bash
# Round-2 closing session on 1 GPU: smoke, GPU suite, the bench line, the e2e ring-depth experiment,
# the ncu launch list of the bench command, one ncu --set full capture of the float-input kernel.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
for n in 3 4 6; do
  echo -n "FMD_HOST_SLOTS=$n: "
  FMD_HOST_SLOTS=$n python bench.py --steps 10 --warmup 3 --no-cpu --no-sublegs --no-sustained 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e', e['value'], 'MS/s', e['gbs_per_direction'], 'GB/s per direction; ceiling', e['ceiling']['gbs_per_direction'], 'frac', e['frac_of_ceiling'])"
done > $O/r02_e2e_ring_depth.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-sustained > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-sustained > $O/ncu_bench.log 2>&1
tail -3 $O/r02_smoke.log; tail -4 $O/r02_pytest_gpu.log; tail -2 $O/r02_bench_n1.err; cat $O/r02_e2e_ring_depth.txt; grep -c fmd_fast $O/r02_launches_bench.csv
