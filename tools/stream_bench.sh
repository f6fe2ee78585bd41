#!/bin/bash
# End-to-end CLI throughput on one long stream (BASELINE configs[4]): file and stdin-pipe input.
# usage: tools/stream_bench.sh [GiB of stream, default 8]
G=${1:-8}
BIN=demodulator_b200/bin/demodulator
REF=oracle/_ref/demodulator_ref
F=/dev/shm/fmd_stream.u8
python - <<PY
import sys; sys.path.insert(0,'.')
from demodulator_b200 import synth
import numpy as np
base = synth.config_signal("cfg3", 64)          # 16 MiB period
with open("$F","wb") as f:
    for _ in range(64): f.write(base.tobytes())  # 1 GiB file
PY
reps=$G
files=""; for i in $(seq $reps); do files="$files $F"; done
echo "stream: $G GiB = $((G*512)) Mi samples, flags -S48000 -l15000 (cfg3/cfg5)"
export FMD_STATS=1
echo "--- file -> /dev/null (1 GiB file)"; $BIN -i $F -o /dev/null -S48000 -l15000
echo "--- file -> file in /dev/shm"; $BIN -i $F -o /dev/shm/fmd_out.f32 -S48000 -l15000; ls -la /dev/shm/fmd_out.f32 | awk "{print \$5}"; rm -f /dev/shm/fmd_out.f32
if [ -x $REF ]; then echo "--- reference CLI, file -> file"; ( time -p $REF -i $F -o /dev/shm/fmd_ref.f32 -S48000 -l15000 ) 2>&1 | grep real; rm -f /dev/shm/fmd_ref.f32; fi
echo "--- stdin pipe (cat x$reps) -> /dev/null"; cat $files | $BIN -i - -o /dev/null -S48000 -l15000
echo "--- stdin pipe, with -L12500 (cfg2 flags)"; cat $files | $BIN -i - -o /dev/null -S96000 -l6500 -L12500
if [ "$(nvidia-smi -L | wc -l)" -gt 1 ]; then
  echo "--- stdin pipe, all GPUs"; cat $files | FMD_DEVICES=all $BIN -i - -o /dev/null -S48000 -l15000
fi
echo "--- cat alone (pipe ceiling)"; /usr/bin/time -f "%e s" cat $files > /dev/null
echo "--- cat | cat > /dev/null"; ( time -p cat $files | cat > /dev/null ) 2>&1 | grep real
if [ -x $REF ]; then
  echo "--- reference CLI, stdin pipe, 1 GiB"; ( time -p cat $F | $REF -i - -o /dev/null -S48000 -l15000 ) 2>&1 | grep real
fi
rm -f $F
