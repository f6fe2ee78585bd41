#!/bin/bash
# quick 1-GPU check: whole GPU suite
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.log
tail -12 $O/r02_pytest_gpu.log
