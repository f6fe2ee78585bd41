"""Deterministic synthetic FM-modulated 8-bit IQ.

phi[n] = 2*pi*cumsum(df*sin(2*pi*fm*n/fs))/fs
I,Q    = clip(rint(127.5 + A*(cos phi, sin phi) + N(0, sigma)), 0, 255) as u8, interleaved.
"""
from __future__ import annotations

import numpy as np

CHUNK_BYTES = 262144  # reference DEFAULT_BUF_SIZE (include/matrix.h:37-39)

# name -> (signal parameters, reference CLI filter options)
CONFIGS = {
    # BASELINE.json configs[0]: 192 kS/s, -S96000 -l12500, no input lowpass
    "cfg1": dict(fs_sig=192e3, df=5e3, fm=1e3, seed=1,
                 opts=dict(fs=96000.0, fc_out=12500.0, fc_in=0.0)),
    # configs[1]: -L12500 -l6500 -S96000 (both filter stages)
    "cfg2": dict(fs_sig=192e3, df=5e3, fm=1e3, seed=2,
                 opts=dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
    # configs[2]: 2.4 MS/s, -S48000 -l15000
    "cfg3": dict(fs_sig=2.4e6, df=75e3, fm=1e3, seed=3,
                 opts=dict(fs=48000.0, fc_out=15000.0, fc_in=0.0)),
}


def fm_iq_u8(n_complex: int, fs_sig: float, df: float, fm: float, seed: int,
             amp: float = 100.0, sigma: float = 1.0, start: int = 0) -> np.ndarray:
    """Return 2*n_complex interleaved u8 I,Q bytes."""
    rng = np.random.default_rng(seed)
    n = np.arange(start, start + n_complex, dtype=np.float64)
    phi = 2.0 * np.pi * np.cumsum(df * np.sin(2.0 * np.pi * fm * n / fs_sig)) / fs_sig
    iq = np.empty(2 * n_complex, np.float64)
    iq[0::2] = 127.5 + amp * np.cos(phi)
    iq[1::2] = 127.5 + amp * np.sin(phi)
    iq += rng.normal(0.0, sigma, iq.size)
    return np.clip(np.rint(iq), 0, 255).astype(np.uint8)


def config_signal(name: str, n_chunks: int, chunk_bytes: int = CHUNK_BYTES, seed: int | None = None,
                  fm: float | None = None) -> np.ndarray:
    c = CONFIGS[name]
    return fm_iq_u8(n_chunks * chunk_bytes // 2, c["fs_sig"], c["df"], fm if fm is not None else c["fm"],
                    c["seed"] if seed is None else seed)
