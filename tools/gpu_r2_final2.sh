#!/bin/bash
# Round-2 closing session on 2 GPUs: the multi-GPU tests and BASELINE configs[4] at full size.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
if [ "$1" != "quality-only" ]; then
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu > $O/r02_pytest_multi_gpu.log 2>&1; echo "pytest rc=$? on $N GPUs" >> $O/r02_pytest_multi_gpu.log
{ timeout 600 python tools/cfg5_stream.py --gsamples 16 --devices all
  timeout 600 python tools/cfg5_stream.py --gsamples 16 --devices 0; } > $O/r02_cfg5_stream_n$N.txt 2>&1
tail -4 $O/r02_pytest_multi_gpu.log; cat $O/r02_cfg5_stream_n$N.txt
fi
{ timeout 600 python tools/cfg5_stream.py --gsamples 4 --devices all --quality
  timeout 600 python tools/cfg5_stream.py --gsamples 4 --devices 0 --quality; } > $O/r02_cfg5_quality_stream_n$N.txt 2>&1
cat $O/r02_cfg5_quality_stream_n$N.txt
