"""CPU: the oracle port (oracle/fm_oracle.c) against the fixtures generated from the reference
itself (tests/golden/make_golden.py) and, where the prebuilt reference library is present,
against the reference's own functions bit for bit."""
from __future__ import annotations

import hashlib

import numpy as np
import pytest

from demodulator_b200 import synth
from oracle import pyoracle as po
from helpers import B, GOLDEN_CASES, golden_input, load_golden


def _same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_matches_golden(name, oracle_port):
    g, cfg = load_golden(name)
    raw = golden_input(g)
    assert hashlib.sha256(raw.tobytes()).hexdigest() == str(g["raw_sha256"]), "synthetic input drifted"
    out = oracle_port.run(raw, cfg)
    o = out.reshape(-1, B // 4)
    assert bool(g["binary_checked"])
    # same libm as the generating container -> bit-exact; otherwise atan2f may differ in the last ulp
    exact = hashlib.sha256(out.tobytes()).hexdigest() == str(g["out_sha256"])
    for got, want, what in ((o[:, :64], g["head"], "head"), (o[:, -64:], g["tail"], "tail"),
                            (out[::97], g["strided"], "strided")):
        assert np.array_equal(np.isnan(got), np.isnan(want)), what
        m = ~np.isnan(want)
        assert np.abs(got[m].astype(np.float64) - want[m]).max(initial=0.0) <= 5e-7 * max(1.0, np.abs(want[m]).max(initial=0.0)), what
    if not exact:
        pytest.xfail("outputs agree to 5e-7 but not bit for bit: different libm atan2f than the fixture's")


@pytest.mark.parametrize("name", ["edge", "edge_L"])
def test_port_matches_edge_fixture(name, oracle_port):
    g, cfg = load_golden(name)
    out = oracle_port.run(g["raw"], cfg)
    assert np.abs(out - g["out_full"]).max() <= 1e-5  # random bytes + input filter: |values| up to ~60


def test_design_table_port_and_product():
    """processFilterOption: 288 designs (4 families x degree 1..8 x 6 cutoffs) float for float."""
    import demodulator_b200 as fm
    t = np.load("tests/golden/design.npz")
    port = po.port()
    for kind, deg, fc, fs, eps, want in zip(t["kind"], t["degree"], t["fc"], t["fs"], t["eps"], t["sos"]):
        rows = (int(deg) + 1) // 2
        a = port.design(int(kind), int(deg), float(fc), float(fs), float(eps))
        b = fm.design_sos(int(kind), int(deg), float(fc), float(fs), float(eps))
        assert _same_bits(a, want[:rows]), ("port", kind, deg, fc, fs)
        assert _same_bits(b, want[:rows]), ("product", kind, deg, fc, fs)


def test_folded_taps_match_survey_a2():
    """Effective stencil of the section sweep."""
    import demodulator_b200 as fm
    cases = [((0, 3, 12500., 96000.), (1.3968215, -0.465274453), -2.0),
             ((0, 3, 6500., 96000.), (2.15487576, -0.657888889), -2.0),
             ((0, 3, 15000., 48000.), (-0.72244072, -0.368045449), -2.0),
             ((0, 3, 12500., 125000.), (1.76004195, -0.545723319), -2.0),
             ((0, 5, 12500., 96000.), (2.37035728, -0.0990790129), -5.0),
             ((1, 3, 12500., 96000.), (2.07917166, -0.798811436), -2.0)]
    for args, want, dcq in cases:
        sos = fm.design_sos(*args)
        taps, dc = fm.fold_taps(sos, True)
        assert dc == 0.0
        assert np.allclose(taps, want, rtol=0, atol=3e-7), (args, taps)
        taps_q, dc_q = fm.fold_taps(sos, False)
        assert np.array_equal(taps, taps_q) and dc_q == dcq


def test_folded_stencil_reproduces_sweep(oracle_port):
    """Away from the head the sweep IS the folded stencil (the fast kernel relies on this)."""
    import demodulator_b200 as fm
    rng = np.random.default_rng(5)
    x = rng.normal(0, 1, 4096).astype(np.float32)
    for kind, deg in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 5), (0, 6), (1, 7), (0, 8)):
        sos = fm.design_sos(kind, deg, 9600., 96000.)
        L = sos.shape[0]
        taps, dc = fm.fold_taps(sos, True)
        y = oracle_port.sweep_real(x, sos)[:x.size]
        xe = np.concatenate([x, np.zeros(8, np.float32)]).astype(np.float64)
        want = dc + sum(float(t) * xe[i:i + x.size] for i, t in enumerate(taps))
        k0 = 2 * L - 1
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(y[k0:] - want[k0:]).max() <= 2e-6 * scale, (kind, deg)
        assert np.all(y[:L] == 0.0)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libdemod_ref.so not present")
def test_port_is_bit_exact_against_reference_library():
    P, R = po.port(), po.ref()
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 256, 2 * B, dtype=np.uint8)
    assert np.array_equal(P.shift(raw), R.shift(raw))
    x = P.shift(raw)
    assert _same_bits(P.discriminate(x), R.discriminate(x))
    for kw in (dict(fs=96000., fc_out=12500.), dict(fs=96000., fc_out=6500., fc_in=12500.),
               dict(fs=48000., fc_out=15000.), dict(fs=96000., fc_out=9600., fc_in=12500., mode=3),
               dict(fs=96000., fc_out=12500., fc_in=12500., deg_out=5),
               dict(fs=96000., fc_out=9600., fc_in=12500., deg_out=4, deg_in=7, mode=2),
               dict(fs=96000., fc_out=9600., deg_out=8, mode=1), dict(fs=125000., fc_out=12500., deg_out=1)):
        cfg = po.Config(**kw)
        assert _same_bits(P.run(raw, cfg), R.run(raw, cfg)), kw
        sig = synth.config_signal("cfg2", 1)
        assert _same_bits(P.run(sig, cfg), R.run(sig, cfg)), kw
    for bs in (512, 4096, B // 2, 2 * B):
        cfg = po.Config(fs=96000., fc_out=6500., fc_in=12500.)
        r = np.resize(raw, 2 * bs)
        assert _same_bits(P.run(r, cfg, chunk_bytes=bs), R.run(r, cfg, chunk_bytes=bs)), bs


def test_reference_semantics_pinned(oracle_port):
    """The five findings of SURVEY.md section 0 as executable facts about the oracle."""
    P = oracle_port
    # F1: shift is -127 with an int8 wrap
    assert list(P.shift(np.array([0, 126, 127, 128, 254, 255], np.uint8))) == [-127, -1, 0, 1, 127, -128]
    # F2: non-overlapping pairs, phase(s0) - phase(s1), signed zeros
    iq = np.array([3, 0, -3, 0, -3, 0, 3, 0, 0, 0, -5, -5, 0, 0, 0, 0], np.float32)
    got = P.discriminate(iq)
    assert np.allclose(got, [-np.pi, np.pi, np.pi, 0.0])
    # F3/F4: N/4 floats per N bytes, no state across chunks, head of every chunk is zero
    raw = synth.config_signal("cfg1", 2)
    cfg = po.Config(fs=96000., fc_out=12500.)
    out = P.run(raw, cfg)
    assert out.size == raw.size // 4
    assert np.all(out.reshape(2, -1)[:, :2] == 0.0)
    assert _same_bits(out[B // 4:], P.run(raw[B:], cfg))
    # F5: with -L the Q lane carries -1 per section and the arena spill seeds outputs 0..5
    outL = P.run(raw, po.Config(fs=96000., fc_out=6500., fc_in=12500.)).reshape(2, -1)
    assert np.all(outL[:, 0] == 0.0) and np.all(outL[:, 1] == -2.0)
    assert np.abs(outL[:, 2:6]).max() > 5.0 and np.abs(outL[:, 6:]).max() < 3.2


def _stencil_model(raw, cfg, to, dco, tin, dcq, shift):
    """numpy statement of what the streaming kernel computes for one chunk: folded stencils with zero
    extension past the chunk (float32 products and sums in float64 order-insensitive form)"""
    s = shift(raw)
    I, Q = s[0::2].astype(np.float64), s[1::2].astype(np.float64)
    nc = I.size

    def st(v, dc):
        ve = np.concatenate([v, np.zeros(8)])
        z = np.full(nc, float(dc))
        for t, c in enumerate(tin):
            z = z + float(c) * ve[t:t + nc]
        return z
    zi, zq = st(I, 0.0), st(Q, dcq)
    a, b, c, d = zi[0::2], zq[0::2], zi[1::2], zq[1::2]
    x = np.arctan2(-a * d + b * c, a * c + b * d)
    xe = np.concatenate([x, np.zeros(8)])
    y = np.full(x.size, float(dco))
    for t, cc in enumerate(to):
        y = y + float(cc) * xe[t:t + x.size]
    return y


@pytest.mark.parametrize("seed", range(24))
def test_conditioners_under_the_input_filter_are_one_stencil(seed, oracle_port):
    """What lets -q2 / -q3 stream under -L, checked without a GPU against
    the oracle on random designs up to degree 8:
      -q3: the normalised samples through the folded input stencil with the FLOAT 0.0 as the zero
           extension past the chunk;
      -q2: highpass and lowpass stencils composed into one (taps convolved, constants added) --
           including the last outputs of the chunk, where the lowpass reads what the highpass sweep
           spilled past its output region: with at most 3 lowpass taps it only reads the first two
           spilled values, which are fully swept, i.e. what the composed stencil gives on zeros."""
    import demodulator_b200.api as api
    rng = np.random.default_rng(9000 + seed)
    q = 2 + seed % 2
    deg_out, deg_in = int(rng.integers(1, 9)), int(rng.integers(1, 9))
    mode = int(rng.integers(0, 4)) | (q << 2)
    fs = float(rng.choice([48000.0, 96000.0, 250000.0]))
    fc_out, fc_in = float(rng.uniform(0.04, 0.4) * fs), float(rng.uniform(0.05, 0.4) * fs)
    raw = synth.config_signal("cfg2", 1, seed=seed) if seed % 3 else rng.integers(0, 256, B, dtype=np.uint8)
    cfg = po.Config(fs=fs, fc_in=fc_in, fc_out=fc_out, eps=0.2, deg_in=deg_in, deg_out=deg_out, mode=mode)
    want = oracle_port.run(raw, cfg)
    if not np.isfinite(want).all():
        pytest.skip("design blows up in the reference too")
    to, dco = api.fold_taps(api.design_sos(mode & 1, deg_out, fc_out, fs, 0.2), True)
    sos_i = api.design_sos((mode >> 1) & 1, deg_in, fc_in, fs, 0.2)
    ti, dci = api.fold_taps(sos_i, True)
    _, dcq = api.fold_taps(sos_i, False)
    assert dci == 0.0 and len(ti) <= 3
    if q == 3:
        y = _stencil_model(raw, cfg, to, dco, ti, dcq, oracle_port.normalize)
    else:
        sos_h = api.design_sos(2, 3, 1.0, fs, 0.0)  # matrix.c:151
        hi, dhi = api.fold_taps(sos_h, True)
        _, dhq = api.fold_taps(sos_h, False)
        assert dhi == 0.0
        comp = np.convolve(ti.astype(np.float64), hi.astype(np.float64))
        y = _stencil_model(raw, cfg, to, dco, comp, float(dcq) + float(dhq) * float(ti.astype(np.float64).sum()),
                           oracle_port.shift)
    head = 16  # chunk heads are the exact path's (arena garbage with -L)
    scale = max(1.0, float(np.abs(want[head:]).max()))
    err = float(np.abs(y[head:] - want[head:]).max()) / scale
    assert err <= 3e-5, (q, deg_in, deg_out, mode, err)
    tail_err = float(np.abs(y[-16:] - want[-16:]).max()) / scale
    assert tail_err <= 3e-5, ("chunk tail", q, deg_in, deg_out, tail_err)
