#!/bin/bash
# Round-2 GPU-box session (1 GPU): A/B of the staging variants (per-lane cp.async ring = product,
# TMA bulk-copy ring, ring depth 4), burst and sustained (3 s loops), parity of the TMA variant,
# and one ncu --set full capture of the product's cfg2 kernel.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
FMD_LIB=$PWD/variants/tma/libfmdemod.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r02_pytest_tma.log 2>&1; echo "pytest(tma) rc=$?" >> $O/r02_pytest_tma.log
{
  for rep in 1 2; do
  for d in default variants/tma/ variants/stages4/; do
    if [ "$d" = default ]; then unset FMD_LIB; name=default; else export FMD_LIB=$PWD/${d}libfmdemod.so; name=$(basename $d); fi
    for c in cfg2 cfg1; do
      echo -n "$name burst: "; timeout 120 python tools/kbench.py $c 2048 40 2>&1 | tail -1
      echo -n "$name sustained 4400 launches: "; timeout 120 python tools/kbench.py $c 2048 4400 2>&1 | tail -1
    done
  done; done
  unset FMD_LIB
} > $O/r02_kernel_ab2.txt 2>&1
python tools/kbench.py cfg2 2048 5 > $O/plain_kbench.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fmd_fast_kernel -s 3 -c 1 -f -o $O/r02_fast_cfg2 \
    python tools/kbench.py cfg2 2048 5 > $O/ncu_kbench.log 2>&1
ncu -i $O/r02_fast_cfg2.ncu-rep --page raw --csv > $O/r02_fast_cfg2_raw.csv 2>/dev/null
ncu -i $O/r02_fast_cfg2.ncu-rep --page details --csv > $O/r02_fast_cfg2_details.csv 2>/dev/null
ncu -i $O/r02_fast_cfg2.ncu-rep --page source --csv > $O/r02_fast_cfg2_source.csv 2>/dev/null
# the narrow-filter quality case that the segment test first used: is its error the stream's own?
python - <<'PY' > $O/r02_quality_narrow_diag.txt 2>&1
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import demodulator_b200 as fm
from demodulator_b200 import synth
from oracle import pyoracle as po
cfg = po.Config(fs=48000.0, fc_out=4000.0, fc_in=9000.0, deg_out=5, deg_in=4, mode=1)
raw = synth.config_signal("cfg3", 40)
want = po.port().quality(cfg, 2.4e6).run(raw)
with fm.Demodulator(sampleRate=48000.0, lowpassOut=4000.0, lowpassIn=9000.0, outFilterDegree=5, inFilterDegree=4, mode=1, quality=True, inputRate=2.4e6) as d:
    out = d.process(raw)
    print("warm-up samples", d.warmup_samples())
print("one carried-state stream on one GPU, fc/rate = 0.0017, Chebyshev degree 5: max-abs vs serial double oracle", float(np.abs(out - want).max()))
PY
tail -3 $O/r02_pytest_tma.log; cat $O/r02_kernel_ab2.txt; cat $O/r02_quality_narrow_diag.txt
