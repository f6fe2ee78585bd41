#!/bin/bash
# last session of the round: whole GPU suite, accuracy report, -q2 timings -> gpurun_out/
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/accuracy_report.py > gpurun_out/accuracy.log 2>&1; grep -E "^q1|^q2|^quality" gpurun_out/accuracy.log
python - <<'PY' > gpurun_out/q2_timing.txt 2>&1
import sys; sys.path.insert(0, '.')
import torch, numpy as np
import demodulator_b200 as fm
from demodulator_b200 import synth
base = synth.config_signal("cfg1", 16)
for what, nbytes, kw in (("-q2 -S96000 -l12500 (streaming kernel, highpass folded into the input stencil)", 1 << 30, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=8)),
                         ("-q2 -S96000 -L12500 -l6500 (op-for-op path)", 64 << 20, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=8)),
                         ("-q1 -S96000 -l12500 (offset scan + op-for-op path)", 256 << 20, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=4)),
                         ("-q1 -S96000 -L12500 -l6500 (offset scan + op-for-op path)", 64 << 20, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=4)),
                         ("-q3 -S96000 -l12500 (streaming kernel)", 1 << 30, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=12)),
                         ("-q3 -S96000 -L12500 -l6500 (op-for-op path)", 64 << 20, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=12))):
    d_in = torch.from_numpy(base).cuda().repeat((nbytes + base.size - 1) // base.size)[:nbytes].contiguous()
    d_out = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    with fm.Demodulator(**kw) as d:
        s = torch.cuda.current_stream().cuda_stream
        d.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, 2)
        ms = d.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, 5) / 5
    print(f"{what}: {nbytes >> 21} Mi samples in {ms:.3f} ms = {nbytes / 2 / ms / 1e6:.1f} GS/s")
PY
cat gpurun_out/q2_timing.txt
