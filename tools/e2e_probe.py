#!/usr/bin/env python
"""1 GPU: fmd_demod_host against fmd_copy_host (the same copies, no kernels) on pinned host buffers,
alternating, for the slice geometry the environment selects (FMD_HOST_SLICE_MB, FMD_HOST_RAMP,
FMD_HOST_SLOTS).  Prints GB/s per direction for every leg; the product default is what bench.py measures."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demodulator_b200.api import Demodulator, PinnedBuffer  # noqa: E402


def main() -> None:
    nbytes = 2 << 30
    steps = int(os.environ.get("PROBE_STEPS", "5"))
    dem = Demodulator(sampleRate=96000.0, lowpassIn=12500.0, lowpassOut=6500.0, inFilterDegree=3, outFilterDegree=3)
    pin_in, pin_out = PinnedBuffer(nbytes), PinnedBuffer(nbytes)
    rng = np.random.default_rng(5)
    block = rng.integers(0, 256, 1 << 24, dtype=np.uint8)
    for o in range(0, nbytes, block.size):
        pin_in.u8[o:o + block.size] = block
    pin_out.u8[:] = 0

    def leg(fn):
        fn(pin_in.ptr, nbytes, pin_out.ptr)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(pin_in.ptr, nbytes, pin_out.ptr)
        return nbytes * steps / (time.perf_counter() - t0) / 1e9

    tag = " ".join(f"{k}={os.environ[k]}" for k in ("FMD_HOST_SLICE_MB", "FMD_HOST_RAMP", "FMD_HOST_SLOTS") if k in os.environ) or "default"
    order = os.environ.get("PROBE_ORDER", "dcdc")
    res = []
    for ch in order:
        res.append((ch, leg(dem.process_ptr if ch == "d" else dem.copy_ptr)))
    d = [v for k, v in res if k == "d"]
    c = [v for k, v in res if k == "c"]
    print(f"{tag:48s} order {order}: demod " + " ".join(f"{v:5.1f}" for v in d) + "  copy-only " + " ".join(f"{v:5.1f}" for v in c)
          + f"  GB/s per direction; best ratio {max(d) / max(c):.3f}", flush=True)
    pin_in.free(); pin_out.free(); dem.close()


if __name__ == "__main__":
    main()
