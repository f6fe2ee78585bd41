#!/usr/bin/env python
"""tools/qbench.py -- the opt-in quality pipeline (FMD_FLAG_QUALITY) on B200: one JSON line per run,
shaped like bench.py's.

    python tools/qbench.py [--config quality_cfg3] [--samples-mi 256] [--steps 10] [--warmup 3] [--no-cpu]

A step = one call of fmd_demod_device over one device-resident batch of `samples` IQ samples of ONE
stream (a quality context carries filter state, so a batch is one stream, not many channels):
[input-cascade aggregate + scan] + main kernel + output scan + zero-input-response kernel.
Algorithmic bytes per sample: 2 read + 4/M written (M = input rate / output rate).
`e2e` = fmd_demod_host on pinned host buffers; `cpu_baseline` = the serial double-precision oracle
(oracle/fm_quality_oracle.c, one core) -- the reference has no such mode to time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

CASES = {
    "quality_m1": dict(sig="cfg1", rate=192e3, kw=dict(sampleRate=192000.0, lowpassOut=20000.0),
                       what="192 kS/s IQ -> -Q -r192000 -S192000 -l20000 (M=1: no decimation, 4 B out per sample)"),
    "quality_cfg1": dict(sig="cfg1", rate=192e3, kw=dict(sampleRate=96000.0, lowpassOut=12500.0),
                         what="192 kS/s IQ -> -Q -r192000 -S96000 -l12500 (M=2, output cascade only)"),
    "quality_cfg2": dict(sig="cfg2", rate=192e3, kw=dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0),
                         what="192 kS/s IQ -> -Q -r192000 -S96000 -L12500 -l6500 (M=2, both cascades)"),
    "quality_cfg3": dict(sig="cfg3", rate=2.4e6, kw=dict(sampleRate=48000.0, lowpassOut=15000.0),
                         what="2.4 MS/s IQ -> -Q -r2400000 -S48000 -l15000 (M=50, fused decimation)"),
}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="quality_cfg3", choices=sorted(CASES))
    ap.add_argument("--samples-mi", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident leg only (clean ncu launch lists)")
    a = ap.parse_args()

    import torch
    import bench
    import demodulator_b200 as fm
    from demodulator_b200 import synth
    from demodulator_b200.api import PinnedBuffer

    case = CASES[a.config]
    sig = synth.CONFIGS[case["sig"]]
    n = a.samples_mi << 20
    base = synth.fm_iq_u8(1 << 21, sig["fs_sig"], sig["df"], sig["fm"], 100)
    d_in = torch.from_numpy(base).cuda().repeat((2 * n + base.size - 1) // base.size)[:2 * n].contiguous()
    dem = fm.Demodulator(quality=True, inputRate=case["rate"], **case["kw"])
    decim = int(round(case["rate"] / case["kw"]["sampleRate"]))
    d_out = torch.empty(dem.max_out_count(2 * n) + 8, dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    sampler = bench.ClockSampler(0)
    with torch.cuda.stream(stream):
        for _ in range(max(a.warmup, 3)):
            dem.process_device(d_in.data_ptr(), 2 * n, d_out.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        sampler.start()
        dem.profile_kernels(True)
        dem.kernel_time()
        l0 = dem.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(a.steps):
            dem.process_device(d_in.data_ptr(), 2 * n, d_out.data_ptr(), stream.cuda_stream)
        ev1.record(stream)
        stream.synchronize()
        ms = ev0.elapsed_time(ev1)
        k_ms, k_n = dem.kernel_time()
        dem.profile_kernels(False)
        launches = dem.launches - l0
    value = n * a.steps / (ms * 1e-3) / 1e6

    e2e_n = min(n, 256 << 20)
    e2e_steps = 3
    got = 0
    e2e_s = float("inf")
    if not a.no_e2e:
        pin_in = PinnedBuffer(2 * e2e_n)
        pin_out = PinnedBuffer(4 * (dem.max_out_count(2 * e2e_n) + 8))
        torch.from_numpy(pin_in.u8).copy_(d_in[:2 * e2e_n])
        dem.process_ptr(pin_in.ptr, 2 * e2e_n, pin_out.ptr)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            got = dem.process_ptr(pin_in.ptr, 2 * e2e_n, pin_out.ptr)
        e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    cpu = None
    if not a.no_cpu:
        from oracle import pyoracle as po
        kw = case["kw"]
        cfg = po.Config(fs=kw["sampleRate"], fc_out=kw["lowpassOut"], fc_in=kw.get("lowpassIn", 0.0))
        m = 16 << 20
        raw = d_in[:2 * m].cpu().numpy()
        q = po.port().quality(cfg, case["rate"])
        t0 = time.perf_counter()
        want = q.run(raw)
        sec = time.perf_counter() - t0
        dem.reset()
        out = dem.process(raw)
        err = float(np.abs(out.astype(np.float64) - want).max())
        cpu = {"value": round(m / sec / 1e6, 2), "unit": "MS/s", "cores": 1, "kind": "port",
               "sample": f"{m >> 20} Mi samples of the same stream, oracle/fm_quality_oracle.c (serial, double)",
               "max_abs_err_vs_gpu": err}

    peak, peak_src = bench.measured_peak_gbs()
    traffic = None
    try:  # dram bytes of the main kernel per launch from the committed ncu capture, scaled to this launch size
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(a.config)
        if e:
            traffic = float(e["dram_bytes_per_launch"]) * n / float(e["samples_per_launch"])
    except Exception:
        pass
    bps = 2.0 + 4.0 / decim
    k_avg = k_ms / max(k_n, 1)
    line = {
        "metric": "input MS/s (complex IQ samples/s), quality pipeline, 1 GPU", "value": round(value, 1), "unit": "MS/s",
        "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 4),
        "higher_is_better": True, "dtype": "f32", "data": "synthetic",
        "config": {"workload": case["what"], "name": a.config, "samples_per_step": n, "decimation": decim,
                   "l2": f"{2 * n >> 20} MiB of input per step vs 126 MB L2"},
        "roofline": {"bound": "hbm", "achieved": round(bps * n / (ms / a.steps * 1e-3) / 1e9, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(bps * n / (ms / a.steps * 1e-3) / 1e9 / peak, 4), "traffic": traffic,
                     "algorithmic_bytes_per_launch": bps * n,
                     "algorithmic_bytes_per_sample": bps, "peak_source": peak_src,
                     "kernel": "fmq_main_kernel", "kernel_ms_avg": round(k_avg, 4),
                     "kernel_share_of_step": round(k_ms / max(ms, 1e-9), 4),
                     "note": "whole step (all launches); the pipeline is issue-bound, not HBM-bound: one atan2 "
                             "and one or three recursive cascades per input sample"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_n * e2e_steps / e2e_s / 1e6, 1), "unit": "MS/s", "h2d_bytes_per_step": 2 * e2e_n,
                "d2h_bytes_per_step": 4 * got, "api": "fmd_demod_host (pinned host buffers)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
