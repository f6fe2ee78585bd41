"""numpy statement of the fused -q1 kernel's tile logic (demodulator_b200/csrc/fmd_q1_fused.cuh) against
the oracle, without a GPU: which samples a tile conditions, what it keeps past its two runs (offsets
walked on past the front run, walked BACKWARDS before the back run, the chunk's last tile running into
its own back run, the zero extension), and where its outputs land."""
import numpy as np
import pytest

from demodulator_b200 import api, synth
from oracle import pyoracle as po

TS = 2048  # steps per tile (FMQ1_TILE)
B = synth.CHUNK_BYTES


def offsets_serial(raw, esr):
    """the reference's float32 offsets before every step of one chunk (matrix.c:129-138), from 0"""
    u = raw.astype(np.float32)
    I, Q = u[0::2], u[1::2]
    nc = I.size
    steps = nc // 2
    off = np.zeros((steps + 1, 2), np.float32)
    o0 = o1 = np.float32(0)
    e = np.float32(esr)
    for k in range(steps):
        off[k] = (o0, o1)
        fi, bi = I[k] - o0, I[nc - 1 - k] - o0
        fq, bq = Q[k] - o1, Q[nc - 1 - k] - o1
        o0 = np.float32(o0 + np.float32(np.float32(fi + bi) * e))
        o1 = np.float32(o1 + np.float32(np.float32(fq + bq) * e))
    off[steps] = (o0, o1)
    return off


def fused_model(raw, fs, tin, dcq, tout, dco, head):
    u = raw.astype(np.float32)
    I, Q = u[0::2], u[1::2]
    nc = I.size
    steps = nc // 2
    esr = np.float32(50.0 / fs)
    a = np.float32(1.0 - 2.0 * float(esr))
    off = offsets_serial(raw, esr)
    nti, nto = len(tin), len(tout)
    LA = 2 * (nto - 1) + (nti - 1 if nti > 1 else 0)
    out = np.full(nc // 2, np.nan, np.float64)
    for s0 in range(0, steps, TS):
        k = np.arange(s0, s0 + TS)
        cf = np.zeros((TS + LA, 2))
        cb = np.zeros((TS + LA, 2))
        cf[:TS, 0], cf[:TS, 1] = I[k] - off[k, 0], Q[k] - off[k, 1]
        # back run in address order: index i <-> complex nc - s0 - TS + i <-> step s0 + TS - 1 - i
        kb = s0 + TS - 1 - np.arange(TS)
        cb[:TS, 0], cb[:TS, 1] = I[nc - 1 - kb] - off[kb, 0], Q[nc - 1 - kb] - off[kb, 1]
        last_tile = s0 + TS == steps
        if not last_tile:  # offsets walked on from the tile's end values
            kk = s0 + TS + np.arange(LA)
            cf[TS:, 0], cf[TS:, 1] = I[kk] - off[kk, 0], Q[kk] - off[kk, 1]
        if s0 > 0:  # offsets walked backwards from the tile's start values
            o = off[s0].astype(np.float32).copy()
            for i in range(LA):
                kq = s0 - 1 - i
                o[0] = np.float32(np.float32(o[0] - esr * np.float32(I[kq] + I[nc - 1 - kq])) * np.float32(1.0 / a))
                o[1] = np.float32(np.float32(o[1] - esr * np.float32(Q[kq] + Q[nc - 1 - kq])) * np.float32(1.0 / a))
                cb[TS + i] = (I[nc - 1 - kq] - o[0], Q[nc - 1 - kq] - o[1])
        if last_tile:
            cf[TS:] = cb[:LA]

        def x_of(run, n):
            kx = np.arange(n)
            zr = [None, None]
            zi = [None, None]
            for w in range(2):
                c = 2 * kx + w
                r, im = run[c, 0].copy(), run[c, 1].copy()
                if nti > 0:
                    im = im + dcq
                    for t in range(1, nti):
                        r = r + tin[t] * run[c + t, 0]
                        im = im + tin[t] * run[c + t, 1]
                zr[w], zi[w] = r, im
            re = zr[0] * zr[1] + zi[0] * zi[1]
            imag = -zr[0] * zi[1] + zi[0] * zr[1]
            return np.arctan2(imag, re)
        nx = TS // 2 + nto - 1
        xf, xb = x_of(cf, nx), x_of(cb, nx)
        if s0 == 0:
            xb[TS // 2:] = 0.0
        for run_x, first in ((xf, s0 // 2), (xb, (nc - s0 - TS) // 2)):
            y = np.full(TS // 2, float(dco))
            for t in range(nto):
                y = y + float(tout[t]) * run_x[t:t + TS // 2]
            out[first:first + TS // 2] = y
    out[:head] = np.nan  # the op-for-op path's
    return out


@pytest.mark.parametrize("case", [dict(fs=96000.0, fc_out=12500.0), dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0),
                                  dict(fs=48000.0, fc_out=15000.0, fc_in=9000.0, deg_out=5, deg_in=5, mode=3),
                                  dict(fs=250000.0, fc_out=30000.0, deg_out=8, mode=1)])
def test_fused_q1_tile_logic_matches_the_oracle(case, oracle_port):
    cfg = po.Config(mode=case.get("mode", 0) | (1 << 2), **{k: v for k, v in case.items() if k != "mode"})
    raw = synth.config_signal("cfg2", 1, seed=7)
    want = oracle_port.run(raw, cfg)  # first chunk of a stream: the static offsets start at 0
    tout, dco = api.fold_taps(api.design_sos(cfg.mode & 1, cfg.deg_out, cfg.fc_out, cfg.fs, cfg.eps), True)
    tin, dcq = np.zeros(0, np.float32), 0.0
    if cfg.fc_in:
        sos_i = api.design_sos((cfg.mode >> 1) & 1, cfg.deg_in or cfg.deg_out, cfg.fc_in, cfg.fs, cfg.eps)
        ti, dci = api.fold_taps(sos_i, True)
        _, dq = api.fold_taps(sos_i, False)
        assert dci == 0.0
        tin, dcq = ti.astype(np.float64) / float(ti[0]), float(dq) / float(ti[0])
    head = 16
    got = fused_model(raw, cfg.fs, tin, dcq, tout.astype(np.float64), float(dco), head)
    assert np.isfinite(got[head:]).all(), "an output was never produced"
    scale = max(1.0, float(np.abs(want[head:]).max()))
    err = float(np.abs(got[head:] - want[head:]).max())
    assert err <= 2e-5 * scale, (case, err)
