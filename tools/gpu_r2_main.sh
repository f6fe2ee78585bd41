#!/bin/bash
# Round-2 GPU-box session (1 GPU): smoke, whole GPU suite, the bench line (all legs), conditioner
# timings, the host's -march=native resolution.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r02_smoke.log
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
python tools/conditioner_timing.py > $O/r02_conditioner_timing.txt 2>&1
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?" >> $O/r02_bench_n1.err
{ gcc -march=native -Q --help=target 2>/dev/null | grep -E '^\s+-m(arch|tune)='; grep -m1 "model name" /proc/cpuinfo; nproc; } > $O/r02_native_march.txt 2>&1
tail -4 $O/r02_smoke.log; tail -15 $O/r02_pytest_gpu.log; cat $O/r02_conditioner_timing.txt; tail -3 $O/r02_bench_n1.err; cut -c1-1500 $O/r02_bench_n1.json
