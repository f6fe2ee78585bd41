#!/bin/bash
# 1-GPU check: whole GPU suite + conditioner timings
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
timeout 300 python tools/conditioner_timing.py > $O/r02_conditioner_timing.txt 2>&1
tail -12 $O/r02_pytest_gpu.log; cat $O/r02_conditioner_timing.txt
