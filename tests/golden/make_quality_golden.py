#!/usr/bin/env python
"""Generate tests/golden/quality_*.npz with scipy (authoring container; scipy is not needed to
USE the fixtures).

The quality pipeline (FMD_FLAG_QUALITY) has no counterpart in the
reference, so its oracle (oracle/fm_quality_oracle.c) is pinned against an independent
implementation instead: scipy.signal.sosfilt (float64, direct form II transposed) over the
section rows the reference's design code produces at the input rate, numpy's angle() as the
discriminator, a plain stride as the decimator.

Stored per case: the input recipe, the section rows, the head, the tail and a strided subset of
the decimated output (float32), and strided samples of the undecimated intermediate signals.

Usage:  python tests/golden/make_quality_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.signal as sg

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from demodulator_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, signal config, complex samples, input rate, Config kwargs (fs = output rate)
CASES = [
    ("quality_cfg1", "cfg1", 3 * 131072, 192e3, dict(fs=96000.0, fc_out=12500.0)),
    ("quality_cfg2", "cfg2", 3 * 131072, 192e3, dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
    ("quality_cfg3", "cfg3", 3 * 131072 + 777, 2.4e6, dict(fs=48000.0, fc_out=15000.0)),
    ("quality_d5_M1", "cfg2", 100001, 192e3, dict(fs=192000.0, fc_out=20000.0, fc_in=30000.0, deg_out=5, deg_in=2)),
    ("quality_m3_d6", "cfg1", 70000, 192e3, dict(fs=48000.0, fc_out=9600.0, fc_in=25000.0, deg_out=6, deg_in=4, mode=3)),
]


def run_scipy(raw: np.ndarray, sos_in, sos_out, decim: int):
    s = raw.astype(np.float64) - 127.5
    zi, zq = s[0::2], s[1::2]
    if sos_in is not None:
        zi, zq = sg.sosfilt(sos_in, zi), sg.sosfilt(sos_in, zq)
    z = zi + 1j * zq
    x = np.angle(z * np.conj(np.concatenate([[0.0], z[:-1]])))
    v = sg.sosfilt(sos_out, x)
    return z, x, v, v[::decim].astype(np.float32)


def main() -> None:
    port = po.port()
    for name, sig, n, rate, kw in CASES:
        cfg = po.Config(**kw)
        decim = int(round(rate / cfg.fs))
        c = synth.CONFIGS[sig]
        raw = synth.fm_iq_u8(n, c["fs_sig"], c["df"], c["fm"], c["seed"])
        sos_out = port.design(cfg.mode & 1, cfg.deg_out, cfg.fc_out, rate, cfg.eps)
        deg_in = cfg.deg_in or cfg.deg_out
        sos_in = port.design((cfg.mode >> 1) & 1, deg_in, cfg.fc_in, rate, cfg.eps) if cfg.fc_in else None
        z, x, v, y = run_scipy(raw, None if sos_in is None else sos_in.astype(np.float64),
                               sos_out.astype(np.float64), decim)
        stride = 97
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"), sig_kind=sig, n=n, input_rate=rate, decim=decim,
            cfg=np.array([cfg.fs, cfg.fc_in, cfg.fc_out, cfg.eps, cfg.deg_in, cfg.deg_out, cfg.mode], np.float64),
            sos_out=sos_out, sos_in=np.zeros((0, 6), np.float32) if sos_in is None else sos_in,
            n_out=y.size, y_head=y[:4096], y_tail=y[-2048:], y_stride=11, y_s=y[::11], stride=stride, z_s=np.stack([z.real[::stride], z.imag[::stride]]).astype(np.float64),
            x_s=x[::stride], v_s=v[::stride])
        print(f"{name}: n={n} M={decim} outputs={y.size} |y|max={np.abs(y).max():.4f}")


if __name__ == "__main__":
    main()
