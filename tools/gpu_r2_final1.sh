#!/bin/bash
# Round-2 closing session on 1 GPU: smoke, GPU suite, the bench line, conditioner timings.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
timeout 300 python tools/conditioner_timing.py > $O/r02_conditioner_timing.txt 2>&1
tail -8 $O/r02_smoke.log; tail -4 $O/r02_pytest_gpu.log; tail -2 $O/r02_bench_n1.err; cut -c1-300 $O/r02_bench_n1.json; cat $O/r02_conditioner_timing.txt
