"""GPU: the opt-in quality pipeline (FMD_FLAG_QUALITY) through the C ABI against its oracle
(oracle/fm_quality_oracle.c, pinned to scipy fixtures by tests/test_quality_oracle.py).

Bar: the same as the reference-compatible path -- decimation indices identical (output m is input
sample m*M of the stream, whatever the call boundaries), floats within max-abs 1e-4."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import demodulator_b200 as fm
from demodulator_b200 import synth
from oracle import pyoracle as po
from helpers import TOL
from test_quality_oracle import QUALITY_CASES, check_against_fixture, load_quality

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")


def quality_kwargs(cfg: po.Config, rate: float) -> dict:
    return dict(sampleRate=cfg.fs, lowpassIn=cfg.fc_in, lowpassOut=cfg.fc_out, epsilon=cfg.eps,
                inFilterDegree=cfg.deg_in, outFilterDegree=cfg.deg_out, mode=cfg.mode, quality=True, inputRate=rate)


def max_err(got: np.ndarray, want: np.ndarray) -> float:
    assert got.shape == want.shape, (got.shape, want.shape)
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) if want.size else 0.0


@pytest.mark.parametrize("name", QUALITY_CASES)
def test_quality_fixture_and_oracle(name, oracle_port):
    g, cfg, raw, rate = load_quality(name)
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        out = d.process(raw)
        launches = d.launches
    check_against_fixture(out, g, TOL, name)
    err = max_err(out, oracle_port.quality(cfg, rate).run(raw))
    print(f"{name}: {out.size} floats, {launches} launches, max-abs vs oracle {err:.2e}")
    assert err <= TOL and launches >= 3


@pytest.mark.parametrize("name,sizes", [
    ("quality_cfg2", (2, 4, 30, 64, 8192, 8190, 20000, 131072 + 6, 100000)),
    ("quality_cfg3", (50, 49, 1, 8191, 8193, 300000)),
    ("quality_d5_M1", (1, 1, 2, 33, 16384, 50000)),
])
def test_quality_stream_continues_across_calls(name, sizes, oracle_port):
    """filter state, the discriminator's neighbour and the decimation phase are carried from call
    to call: any split gives the unsplit result; reset() starts over"""
    g, cfg, raw, rate = load_quality(name)
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        for rep in range(2):
            off = 0
            worst = 0.0
            q = oracle_port.quality(cfg, rate)
            for n in sizes:
                seg = raw[2 * off:2 * (off + n)]
                want = q.run(seg)
                assert d.out_count(seg.size) == want.size  # depends on the stream position
                got = d.process(seg)
                worst = max(worst, max_err(got, want))
                off += n
            assert worst <= TOL, (name, rep, worst)
            d.reset()


def test_quality_every_degree_combination(oracle_port):
    """one kernel instantiation per (input degree 0..6, output degree 1..6), in both thread spans
    between them: Butterworth and Chebyshev, against the oracle, in two calls that split a tile"""
    rate = 192e3
    c = synth.CONFIGS["cfg2"]
    raw = synth.fm_iq_u8(80000, c["fs_sig"], c["df"], c["fm"], 21)
    worst = 0.0
    for deg_in in range(0, 7):
        for deg_out in range(1, 7):
            # M = 4 runs the 64-samples-per-thread kernels, M = 8 the 128-samples-per-thread ones
            fs = 48000.0 if (deg_in + deg_out) % 2 == 0 else 24000.0
            mode = 3 if (deg_in + deg_out) % 3 == 0 else 0
            cfg = po.Config(fs=fs, fc_out=9000.0 + 500.0 * deg_out, fc_in=30000.0 if deg_in else 0.0,
                            deg_in=deg_in, deg_out=deg_out, mode=mode, eps=0.1)
            want = oracle_port.quality(cfg, rate).run(raw)
            with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
                out = np.concatenate([d.process(raw[:2 * 37001]), d.process(raw[2 * 37001:])])
            err = max_err(out, want)
            worst = max(worst, err)
            assert err <= TOL, (deg_in, deg_out, mode, fs, err)
    print(f"42 degree combinations: worst max-abs {worst:.2e}")


def test_quality_edge_geometry(oracle_port):
    """decimation larger than a tile, calls shorter than the decimation, empty calls, odd byte counts"""
    rate = 2.4e6
    cfg = po.Config(fs=100.0, fc_out=15000.0)  # M = 24000 > 16384 samples per tile
    c = synth.CONFIGS["cfg3"]
    raw = synth.fm_iq_u8(120000, c["fs_sig"], c["df"], c["fm"], 22)
    q = oracle_port.quality(cfg, rate)
    assert q.decim == 24000
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        off, got, want = 0, [], []
        for n in (0, 5, 23990, 10, 0, 50000, 45995):
            seg = raw[2 * off:2 * (off + n)]
            w = q.run(seg)
            assert d.out_count(seg.size) == w.size
            got.append(d.process(seg))
            want.append(w)
            off += n
        assert max_err(np.concatenate(got), np.concatenate(want)) <= TOL
        assert np.concatenate(want).size == 5  # samples 0, 24000, ..., 96000
        with pytest.raises(fm.FmdError, match="even number of bytes"):
            d.process(raw[:7])


def test_quality_device_resident_many_tiles(oracle_port):
    """16 Mi samples resident in HBM, one call: 1024 tiles through the tile-level scan"""
    import torch
    cfg, rate = po.Config(fs=48000.0, fc_out=15000.0), 2.4e6
    n = 16 << 20
    c = synth.CONFIGS["cfg3"]
    raw = synth.fm_iq_u8(n, c["fs_sig"], c["df"], c["fm"], 11)
    want = oracle_port.quality(cfg, rate).run(raw)
    d_in = torch.from_numpy(raw).cuda()
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        d_out = torch.full((d.max_out_count(raw.size) + 8,), 7.0, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        d.process_device(d_in.data_ptr(), raw.size, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    assert (out[want.size:] == 7.0).all()  # nothing written past the last output
    err = max_err(out[:want.size], want)
    print(f"device-resident 16 Mi samples: max-abs {err:.2e}")
    assert err <= TOL


def test_quality_with_input_filter_many_tiles(oracle_port):
    """both cascades over 4 Mi samples (256 tiles), in two unequal calls"""
    cfg, rate = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0), 192e3
    n = 4 << 20
    c = synth.CONFIGS["cfg2"]
    raw = synth.fm_iq_u8(n, c["fs_sig"], c["df"], c["fm"], 12)
    want = oracle_port.quality(cfg, rate).run(raw)
    cut = 2 * (3 * 8192 * 100 + 4097)
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        out = np.concatenate([d.process(raw[:cut]), d.process(raw[cut:])])
    err = max_err(out, want)
    print(f"-L quality 4 Mi samples: max-abs {err:.2e}")
    assert err <= TOL


def test_quality_constant_tone_property():
    """a pure tone at f has phase step 2*pi*f/rate per sample: once the filters have settled every
    output equals it (DC gain of the Butterworth cascade is 1), at any size"""
    rate, f, n = 2.4e6, 60e3, 8 << 20
    t = np.arange(n, dtype=np.float64)
    iq = np.empty(2 * n, np.float64)
    iq[0::2] = 127.5 + 100.0 * np.cos(2 * np.pi * f * t / rate)
    iq[1::2] = 127.5 + 100.0 * np.sin(2 * np.pi * f * t / rate)
    raw = np.clip(np.rint(iq), 0, 255).astype(np.uint8)
    with fm.Demodulator(sampleRate=48000.0, lowpassOut=15000.0, quality=True, inputRate=rate) as d:
        out = d.process(raw)
    assert out.size == (n + 49) // 50
    settled = out[1000:]
    assert abs(float(settled.mean()) - 2 * np.pi * f / rate) < 2e-4
    assert float(settled.std()) < 2e-3  # 8-bit quantisation noise through a 15 kHz lowpass


def test_quality_s16_egress(tmp_path_nodash, oracle_port):
    """FMD_FLAG_S16 / -F s16: the decimated outputs as signed 16-bit samples (1.0 = 32768, round to
    nearest, clip) -- the oracle's floats quantised the same way, give or take one step where the GPU's
    float sits on the other side of a rounding boundary"""
    g, cfg, raw, rate = load_quality("quality_cfg3")
    y = oracle_port.quality(cfg, rate).run(raw).astype(np.float64)
    want = np.clip(np.rint(y * 32768.0), -32768, 32767).astype(np.int16)
    with fm.Demodulator(s16=True, **quality_kwargs(cfg, rate)) as d:
        out = np.concatenate([d.process(raw[: 2 * 150001]), d.process(raw[2 * 150001:])])
    assert out.dtype == np.int16 and out.size == want.size
    diff = np.abs(out.astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.05, (int(diff.max()), float((diff != 0).mean()))
    # clipping: a gain-free 0.2 rad signal never clips, so drive the clip path through the CLI on a loud one
    src, dst = tmp_path_nodash / "in.u8", tmp_path_nodash / "out.s16"
    raw.tofile(src)
    r = subprocess.run([CLI, "-i", str(src), "-o", str(dst), "-Q", "-r2400000", "-S48000", "-l15000", "-F", "s16"],
                       capture_output=True, text=True, env=dict(os.environ, FMD_SLOT_MB="1"))
    assert r.returncode == 0, r.stderr
    cli = np.fromfile(dst, np.int16)
    want_cli = want[: oracle_port.quality(cfg, rate).run(raw[: raw.size & ~3]).size]
    assert cli.size == want_cli.size and np.abs(cli.astype(np.int32) - want_cli.astype(np.int32)).max() <= 1
    with pytest.raises(fm.FmdError, match="FMD_FLAG_S16"):
        fm.Demodulator(s16=True)


def test_quality_cli(tmp_path_nodash, oracle_port):
    g, cfg, raw, rate = load_quality("quality_cfg3")
    src, dst = tmp_path_nodash / "in.u8", tmp_path_nodash / "out.f32"
    raw.tofile(src)
    r = subprocess.run([CLI, "-i", str(src), "-o", str(dst), "-Q", "-r2400000", "-S48000", "-l15000"],
                       capture_output=True, text=True, env=dict(os.environ, FMD_SLOT_MB="1"))
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst, np.float32)
    # the CLI drops a ragged tail below 4 bytes (2 samples), like the reference-compatible mode
    want = oracle_port.quality(cfg, rate).run(raw[: raw.size & ~3])
    assert max_err(out, want) <= TOL


@pytest.mark.parametrize("case", [dict(sig="cfg3", rate=2.4e6, fs=48000.0, fc_out=15000.0),
                                  dict(sig="cfg2", rate=192e3, fs=96000.0, fc_out=6500.0, fc_in=12500.0),
                                  dict(sig="cfg3", rate=2.4e6, fs=48000.0, fc_out=14000.0, fc_in=30000.0, deg_out=5, deg_in=4, mode=1)])
def test_quality_segments_with_warm_up_match_the_stream(case, oracle_port):
    """The multi-GPU form of the stream (fmd_submit_segment): independent segments that start
    fmd_warmup_samples early from a ZERO state and drop the prefix's outputs must reproduce the one
    carried-state stream.  Run here on one GPU with segments in scrambled order -- nothing may depend
    on which context, or in which order, a segment is processed."""
    from demodulator_b200.api import PinnedBuffer
    case = dict(case)
    sig, rate = case.pop("sig"), case.pop("rate")
    cfg = po.Config(**case)
    raw = synth.config_signal(sig, 40)
    n = raw.size // 2
    want = oracle_port.quality(cfg, rate).run(raw)
    with fm.Demodulator(**quality_kwargs(cfg, rate)) as d:
        W = d.warmup_samples()
        assert W % 8 == 0 and 0 < W < n // 8
        seg = (n // 5 + 12345) // 8 * 8  # ragged against tiles, decimation and chunks; multiple of 8 samples
        d.ring_create(2, 2 * (seg + W) // 262144 * 262144 + 262144)
        pin = PinnedBuffer(raw.size)
        pin.u8[:] = raw
        outs = PinnedBuffer(4 * (n + 8))
        starts = list(range(0, n, seg))
        offs, counts, o = {}, {}, 0
        for g0 in starts:
            ns = min(seg, n - g0)
            counts[g0] = d.segment_out_count(g0, ns)
            offs[g0] = o
            o += counts[g0]
        assert o == want.size
        order = starts[::-2] + starts[-2::-2] if len(starts) > 1 else starts
        assert sorted(order) == starts
        for i, g0 in enumerate(order):
            ns = min(seg, n - g0)
            warm = min(W, g0)
            d.wait(i % 2)
            d.submit_segment(i % 2, pin.ptr + 2 * (g0 - warm), 2 * warm, 2 * ns, g0, outs.ptr + 4 * offs[g0])
        d.wait(0); d.wait(1)
        got = outs.view(np.float32)[: want.size].copy()
        pin.free(); outs.free()
    err = max_err(got, want)
    print(f"segments with warm-up {W}: max-abs {err:.2e}")
    assert err <= 2e-5, err
