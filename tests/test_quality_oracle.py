"""CPU suite for the opt-in quality pipeline (FMD_FLAG_QUALITY).

The reference computes nothing of the kind, so the chain of evidence is:
  scipy.signal.sosfilt fixtures (tests/golden/quality_*.npz, make_quality_golden.py)
    -> pin the serial double-precision oracle (oracle/fm_quality_oracle.c)
    -> which pins the numpy model of the parallel-scan algebra + the host-built power tables
       (tests/quality_model.py over libfmdemod.so's fmq_test_build)
    -> and, on the GPU box, the kernels themselves (tests/test_quality_gpu.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pytest

import demodulator_b200 as fm
from demodulator_b200 import synth
from oracle import pyoracle as po
import quality_model as qm
from helpers import GOLDEN

QUALITY_CASES = ["quality_cfg1", "quality_cfg2", "quality_cfg3", "quality_d5_M1", "quality_m3_d6"]


def load_quality(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    fs, fc_in, fc_out, eps, deg_in, deg_out, mode = g["cfg"]
    cfg = po.Config(fs=float(fs), fc_in=float(fc_in), fc_out=float(fc_out), eps=float(eps),
                    deg_in=int(deg_in), deg_out=int(deg_out), mode=int(mode))
    c = synth.CONFIGS[str(g["sig_kind"])]
    raw = synth.fm_iq_u8(int(g["n"]), c["fs_sig"], c["df"], c["fm"], c["seed"])
    return g, cfg, raw, float(g["input_rate"])


def check_against_fixture(y: np.ndarray, g, tol: float, what: str) -> float:
    assert y.size == int(g["n_out"]), (what, y.size, int(g["n_out"]))
    err = 0.0
    for got, want in ((y[:4096], g["y_head"]), (y[-2048:], g["y_tail"]), (y[::int(g["y_stride"])], g["y_s"])):
        assert got.shape == want.shape
        err = max(err, float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()))
    assert err <= tol, f"{what}: max-abs {err:.3e} > {tol:.1e}"
    return err


@pytest.mark.parametrize("name", QUALITY_CASES)
def test_oracle_matches_scipy_fixture(name, oracle_port):
    g, cfg, raw, rate = load_quality(name)
    q = oracle_port.quality(cfg, rate)
    assert q.decim == int(g["decim"])
    # the design rows are the reference's own design code evaluated at the input rate
    np.testing.assert_array_equal(oracle_port.design(cfg.mode & 1, cfg.deg_out, cfg.fc_out, rate, cfg.eps), g["sos_out"])
    z, x, v = oracle_port.quality(cfg, rate).stages(raw)
    st = int(g["stride"])
    assert np.abs(z[0::2][::st] - g["z_s"][0]).max() < 1e-9 and np.abs(z[1::2][::st] - g["z_s"][1]).max() < 1e-9
    assert np.abs(x[::st] - g["x_s"]).max() < 1e-9
    assert np.abs(v[::st] - g["v_s"]).max() < 1e-9
    check_against_fixture(q.run(raw), g, 1e-7, name)  # float32 rounding of the stored output


def test_oracle_stream_is_split_invariant(oracle_port):
    """carried state: any split of the stream gives the bytes of the unsplit run"""
    g, cfg, raw, rate = load_quality("quality_cfg2")
    raw = raw[:2 * 50000]
    whole = oracle_port.quality(cfg, rate).run(raw)
    q = oracle_port.quality(cfg, rate)
    parts, off = [], 0
    for n in (1, 2, 3, 4097, 8192, 11, 37694):
        parts.append(q.run(raw[2 * off:2 * (off + n)]))
        off += n
    assert off == 50000
    np.testing.assert_array_equal(np.concatenate(parts), whole)


def test_butterworth_rows_are_the_textbook_filter(oracle_port):
    """the reference's design at (fc, rate) is scipy's butter(): quality mode is a true Butterworth"""
    sg = pytest.importorskip("scipy.signal")
    for deg, fc, rate in ((3, 6500.0, 192e3), (5, 15000.0, 2.4e6), (4, 30000.0, 192e3)):
        ours = oracle_port.design(0, deg, fc, rate).astype(np.float64)
        w, h1 = sg.sosfreqz(ours, worN=512)
        _, h2 = sg.sosfreqz(sg.butter(deg, fc, fs=rate, output="sos"), worN=512)
        assert np.abs(h1 - h2).max() < 2e-5


@pytest.fixture(scope="module")
def product_lib():
    return C.CDLL(fm.lib_path())


@pytest.mark.parametrize("name,sizes,T", [
    ("quality_cfg2", (20000, 8192, 33, 1, 50001), 64),
    ("quality_cfg3", (8192 * 3, 31, 32, 8191, 40000), 128),
    ("quality_cfg3", (16384 * 2, 63, 64, 16383, 30000), 64),
    ("quality_d5_M1", (10000, 7, 16384), 64),
    ("quality_d5_M1", (32768, 127, 128, 5), 128),
    ("quality_m3_d6", (70000,), 64),
])
def test_scan_algebra_and_tables_match_oracle(name, sizes, T, oracle_port, product_lib):
    """the three-level scan with the host-built power tables reproduces the serial recursion,
    across calls of awkward lengths (tile = 8192 samples, thread = 32)"""
    g, cfg, raw, rate = load_quality(name)
    q = oracle_port.quality(cfg, rate)
    sos_out = oracle_port.design(cfg.mode & 1, cfg.deg_out, cfg.fc_out, rate, cfg.eps)
    deg_in = cfg.deg_in or cfg.deg_out
    sos_in = oracle_port.design((cfg.mode >> 1) & 1, deg_in, cfg.fc_in, rate, cfg.eps) if cfg.fc_in else None
    st = qm.Stream(product_lib, sos_in, sos_out, q.decim, T)
    off = 0
    for n in sizes:
        seg = raw[2 * off:2 * (off + n)]
        want, got = q.run(seg), st.run(seg)
        assert want.shape == got.shape
        if want.size:
            assert float(np.abs(want.astype(np.float64) - got).max()) < 5e-6, (name, off, n)
        off += n


def test_quality_configuration_is_validated_before_the_device():
    """argument errors are reported as such even on a machine without a GPU"""
    def create(**kw):
        try:
            fm.Demodulator(quality=True, **kw).close()
        except fm.FmdError as e:
            return str(e)
        return ""
    assert "integer" in create(sampleRate=48000.0, inputRate=100000.0)
    assert "input rate" in create(sampleRate=48000.0)
    assert "degrees up to" in create(sampleRate=48000.0, inputRate=96000.0, outFilterDegree=7)
    assert "-q0" in create(sampleRate=48000.0, inputRate=96000.0, mode=12)
    assert "float32" in create(sampleRate=48000.0, inputRate=2.4e6, lowpassOut=100.0)


@pytest.mark.parametrize("case", [dict(sig="cfg3", rate=2.4e6, fs=48000.0, fc_out=15000.0),
                                  dict(sig="cfg2", rate=192e3, fs=96000.0, fc_out=6500.0, fc_in=12500.0),
                                  dict(sig="cfg1", rate=192e3, fs=96000.0, fc_out=12500.0),
                                  dict(sig="cfg3", rate=2.4e6, fs=48000.0, fc_out=14000.0, fc_in=30000.0, deg_out=5, deg_in=4, mode=1),
                                  dict(sig="cfg2", rate=192e3, fs=192000.0, fc_out=400.0, deg_out=6)])
def test_warm_up_prefix_replaces_the_carried_state(case, oracle_port):
    """The multi-GPU split of a quality stream (fmd_submit_segment) starts a segment from a ZERO
    state `warm-up` samples early and drops the prefix's outputs.  Here, without a GPU: the library's
    warm-up length (fmd_warmup_for_sections, from the slowest pole) is enough for the serial
    double-precision oracle restarted that early to agree with the one carried-state stream."""
    from demodulator_b200 import api, synth
    case = dict(case)
    sig, rate = case.pop("sig"), case.pop("rate")
    cfg = po.Config(**case)
    M = int(round(rate / cfg.fs))
    raw = synth.config_signal(sig, 24)
    want = oracle_port.quality(cfg, rate).run(raw)
    W = api.warmup_for_sections(api.design_sos(cfg.mode & 1, cfg.deg_out, cfg.fc_out, rate, cfg.eps))
    if cfg.fc_in:
        # -L without -D inherits the output degree (src/matrix.c:190-192)
        W += api.warmup_for_sections(api.design_sos((cfg.mode >> 1) & 1, cfg.deg_in or cfg.deg_out, cfg.fc_in, rate, cfg.eps))
    n = raw.size // 2
    assert 0 < W < n // 4
    g0 = (n // 2) // M * M          # payload start
    seg_start = (g0 - W) // M * M   # a multiple of M: a fresh oracle stream then has the stream's decimation phase
    seg = oracle_port.quality(cfg, rate).run(raw[2 * seg_start:])
    got, ref = seg[(g0 - seg_start) // M:], want[g0 // M:]
    assert got.size == ref.size
    cold = oracle_port.quality(cfg, rate).run(raw[2 * g0:])[: 2000 // M + 8]
    assert float(np.abs(got - ref).max()) <= 1e-8, (W, float(np.abs(got - ref).max()))
    assert float(np.abs(cold - ref[: cold.size]).max()) > 1e-4  # without the prefix the start is visibly wrong
