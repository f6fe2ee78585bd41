"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import os

import numpy as np

from demodulator_b200 import synth
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
B = synth.CHUNK_BYTES
TOL = 1e-4  # north-star tolerance: max-abs 1e-4 against the reference's CPU output

GOLDEN_CASES = ["cfg1", "cfg2", "cfg3", "default", "d5_L", "m3_L", "m1_d4", "m1_e5_nan", "D7_d4_L", "q3", "q3_L",
                "q2", "q2_L", "q2_d5_D4_L_m3", "q1", "q1_L", "q1_cfg3"]


def load_golden(name: str):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    fs, fc_in, fc_out, eps, deg_in, deg_out, mode = g["cfg"]
    cfg = po.Config(fs=float(fs), fc_in=float(fc_in), fc_out=float(fc_out), eps=float(eps),
                    deg_in=int(deg_in), deg_out=int(deg_out), mode=int(mode))
    return g, cfg


def golden_input(g) -> np.ndarray:
    kind = str(g["sig_kind"])
    if kind == "edge_bytes":
        return g["raw"]
    return synth.config_signal(kind, int(g["nchunks"]))


def demod_kwargs(cfg: po.Config) -> dict:
    return dict(sampleRate=cfg.fs, lowpassIn=cfg.fc_in, lowpassOut=cfg.fc_out, epsilon=cfg.eps,
                inFilterDegree=cfg.deg_in, outFilterDegree=cfg.deg_out, mode=cfg.mode)


def cli_flags(cfg: po.Config) -> list[str]:
    f = [f"-S{cfg.fs:g}", f"-l{cfg.fc_out:g}", f"-d{cfg.deg_out}", f"-e{cfg.eps * 10:g}"]
    if cfg.fc_in:
        f.append(f"-L{cfg.fc_in:g}")
    if cfg.deg_in:
        f.append(f"-D{cfg.deg_in}")
    if cfg.mode:
        f.append(f"-m{cfg.mode}")
    return f


def check_close(got: np.ndarray, want: np.ndarray, tol: float = TOL, what: str = "") -> float:
    """max-abs comparison that also requires NaNs (the reference's, when its design blows up)
    to sit at the same positions."""
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_w, nan_g = np.isnan(want), np.isnan(got)
    assert np.array_equal(nan_w, nan_g), f"{what}: NaN positions differ"
    if nan_w.all():
        return 0.0
    err = float(np.abs(got[~nan_w].astype(np.float64) - want[~nan_w].astype(np.float64)).max())
    assert err <= tol, f"{what}: max-abs error {err:.3e} > {tol:.1e}"
    return err


def snr_db(got: np.ndarray, want: np.ndarray) -> float:
    d = got.astype(np.float64) - want.astype(np.float64)
    return 10.0 * np.log10((want.astype(np.float64) ** 2).sum() / max((d ** 2).sum(), 1e-300))
