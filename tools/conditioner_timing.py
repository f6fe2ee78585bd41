#!/usr/bin/env python
"""tools/conditioner_timing.py -- device-resident throughput of the -q conditioners with and without
the input filter -> profiles/rNN_conditioner_timing.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import demodulator_b200 as fm
from demodulator_b200 import synth

base = synth.config_signal("cfg1", 16)
ROWS = (("-q2 -S96000 -l12500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=8)),
        ("-q2 -S96000 -L12500 -l6500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=8)),
        ("-q3 -S96000 -l12500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=12)),
        ("-q3 -S96000 -L12500 -l6500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=12)),
        ("-q1 -S96000 -l12500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=12500.0, mode=4)),
        ("-q1 -S96000 -L12500 -l6500", 1 << 30, dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0, mode=4)))
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
for what, nbytes, kw in ROWS:
    if ONLY and not what.startswith(ONLY):
        continue
    d_in = torch.from_numpy(base).cuda().repeat((nbytes + base.size - 1) // base.size)[:nbytes].contiguous()
    d_out = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
    with fm.Demodulator(**kw) as d:
        s = torch.cuda.current_stream().cuda_stream
        d.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, 2)
        l0 = d.launches
        ms = d.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, 5) / 5
        per = (d.launches - l0) // 5
    print(f"{what}: {nbytes >> 21} Mi samples in {ms:.3f} ms = {nbytes / 2 / ms / 1e6:.1f} GS/s "
          f"({4 * (nbytes / 2) / ms / 1e6:.0f} GB/s algorithmic, {per} launch(es) per call)", flush=True)
