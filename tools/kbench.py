#!/usr/bin/env python
"""tools/kbench.py -- time fmd_demod_device on a device-resident buffer of a chosen size.
usage: kbench.py <cfg> <MiB per launch> [reps]   (small sizes stay in L2: compute-only timing)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import demodulator_b200 as fm
from demodulator_b200 import synth
import bench
cfg = sys.argv[1]; mib = float(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
nbytes = int(mib * (1 << 20)) // 262144 * 262144
base = synth.config_signal(cfg, 16)
d_in = torch.from_numpy(base).cuda().repeat((nbytes + base.size - 1) // base.size)[:nbytes].contiguous()
d_out = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda")
dem = fm.Demodulator(**bench.FLAGS[cfg])
s = torch.cuda.current_stream().cuda_stream
dem.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, 5)
ms = dem.time_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), s, reps) / reps
print(f"{cfg} {mib:g} MiB/launch: {ms*1e3:.2f} us/launch  {2*nbytes/ms/1e6:.1f} GB/s algorithmic  ({ms*1e3*256/mib:.1f} us per 256 MiB)")
