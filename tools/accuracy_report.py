#!/usr/bin/env python
"""Max-abs error of the CUDA path against the reference (oracle best) for every golden case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import demodulator_b200 as fm
from oracle import pyoracle as po
from helpers import GOLDEN_CASES, load_golden, golden_input, demod_kwargs
chk = po.best()
print(f"checker: {chk.kind}")
for name in GOLDEN_CASES + ["edge", "edge_L"]:
    g, cfg = load_golden(name)
    raw = golden_input(g)
    want = chk.run(raw, cfg)
    for exact in (False, True):
        with fm.Demodulator(exact=exact, **demod_kwargs(cfg)) as d:
            out = d.process(raw)
        m = ~np.isnan(want)
        if not m.any():
            print(f"{name:10s} {'exact' if exact else 'fast '}: all NaN in the reference; NaN positions equal: {np.array_equal(np.isnan(out), np.isnan(want))}")
            continue
        err = np.abs(out[m].astype(np.float64) - want[m])
        steady = np.ones(out.size, bool); steady.reshape(-1, 65536)[:, :8] = False
        es = np.abs(out[m & steady].astype(np.float64) - want[m & steady])
        if es.size == 0:
            es = err
        snr = 10 * np.log10((want[m].astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-300))
        print(f"{name:10s} {'exact' if exact else 'fast '}: max-abs {err.max():.3e} (steady state {es.max():.3e}, p99.99 {np.quantile(es, 0.9999):.2e})  SNR {snr:6.1f} dB  |out|max {np.abs(want[m]).max():.3g}")

# ---- the opt-in quality pipeline against its serial double-precision oracle (pinned to scipy fixtures) ----
from test_quality_oracle import QUALITY_CASES, load_quality
print("\nquality pipeline (FMD_FLAG_QUALITY) vs oracle/fm_quality_oracle.c:")
for name in QUALITY_CASES:
    g, cfg, raw, rate = load_quality(name)
    want = po.port().quality(cfg, rate).run(raw)
    with fm.Demodulator(sampleRate=cfg.fs, lowpassIn=cfg.fc_in, lowpassOut=cfg.fc_out, epsilon=cfg.eps,
                        inFilterDegree=cfg.deg_in, outFilterDegree=cfg.deg_out, mode=cfg.mode, quality=True,
                        inputRate=rate) as d:
        out = np.concatenate([d.process(raw[: 2 * 100003]), d.process(raw[2 * 100003:])]) if raw.size > 2 * 100003 else d.process(raw)
    err = np.abs(out.astype(np.float64) - want)
    snr = 10 * np.log10((want.astype(np.float64) ** 2).sum() / max((err ** 2).sum(), 1e-300))
    print(f"{name:14s}: {out.size:7d} floats (M = {int(g['decim'])}), max-abs {err.max():.3e}, p99.99 {np.quantile(err, 0.9999):.2e}, "
          f"SNR {snr:6.1f} dB, |out|max {np.abs(want).max():.3g}")
