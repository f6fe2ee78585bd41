"""GPU, two or more devices: the path shards with no exchange step -- channels and
chunk-aligned ranges of one stream are independent -- so every multi-GPU arrangement must give
exactly what one GPU gives.  Skipped on a one-GPU box (the driver's GPU tier); run under
`gpurun --gpus 2` (tools/gpu_r2_multi.sh)."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

import demodulator_b200 as fm
from demodulator_b200 import sharding, synth
from demodulator_b200.api import device_count
from oracle import pyoracle as po
from helpers import B, check_close, demod_kwargs

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")


def _ndev() -> int:
    try:
        return device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")


@pytest.fixture
def scratch():
    import shutil
    import tempfile
    from pathlib import Path
    d = tempfile.mkdtemp(prefix="fmdmg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    assert "-" not in d
    yield Path(d)
    shutil.rmtree(d, ignore_errors=True)


@needs2
@pytest.mark.parametrize("flags,kw", [(["-S96000", "-l6500", "-L12500"], dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
                                      (["-S48000", "-l15000"], dict(fs=48000.0, fc_out=15000.0))])
def test_cli_one_stream_over_all_gpus_keeps_order_and_parity(scratch, flags, kw, oracle_best):
    """FMD_DEVICES=all: the slots of ONE stream round-robin over the GPUs and come back in order
    (the reference's reader loop, src/main.c:72-91, sees one ordered stream)."""
    n_chunks = 67  # not a multiple of anything: ragged last slot, every device gets several slots
    raw = synth.config_signal("cfg2", n_chunks)
    raw = np.concatenate([raw, raw[:1000 * 4 + 2]])  # ragged tail
    src = scratch / "in.u8"
    raw.tofile(src)
    outs = {}
    for devs in ("0", "all"):
        dst = scratch / f"out_{devs}.f32"
        env = dict(os.environ, FMD_DEVICES=devs, FMD_SLOT_MB="1", FMD_SLOTS="2", FMD_IO_THREADS="8", FMD_STATS="1")
        r = subprocess.run([CLI, "-i", str(src), "-o", str(dst)] + flags, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        if devs == "all":
            assert f"{_ndev()} GPU(s)" in r.stderr, r.stderr
        outs[devs] = np.fromfile(dst, np.float32)
    assert outs["all"].size == raw.size // 4
    assert np.array_equal(outs["0"].view(np.uint32), outs["all"].view(np.uint32)), "multi-GPU output differs from one GPU"
    whole = n_chunks * B
    check_close(outs["all"][: whole // 4], oracle_best.run(raw[:whole], po.Config(**kw)), what="FMD_DEVICES=all")


@needs2
def test_cli_stdin_stream_over_all_gpus(scratch, oracle_best):
    raw = synth.config_signal("cfg3", 40)
    env = dict(os.environ, FMD_DEVICES="all", FMD_SLOT_MB="1")
    r = subprocess.run([CLI, "-i", "-", "-o", "-", "-S48000", "-l15000"], input=raw.tobytes(), capture_output=True, env=env)
    assert r.returncode == 0, r.stderr
    out = np.frombuffer(r.stdout, np.float32)
    check_close(out, oracle_best.run(raw, po.Config(fs=48000.0, fc_out=15000.0)), what="stdin over all GPUs")


@needs2
def test_channels_sharded_over_devices_match_one_device(oracle_best):
    """BASELINE configs[3] in small: channel c runs on GPU c % N (sharding.assign_channels); every
    channel must come out exactly as on GPU 0 and match the reference."""
    n = _ndev()
    cfg = po.Config(fs=48000.0, fc_out=15000.0)
    chans = [synth.config_signal("cfg3", 3, seed=100 + c, fm=1000.0 + 50.0 * c) for c in range(2 * n)]
    with fm.Demodulator(device=0, **demod_kwargs(cfg)) as d0:
        base = [d0.process(x).copy() for x in chans]
    for rank in range(n):
        with fm.Demodulator(device=rank, **demod_kwargs(cfg)) as d:
            for c in sharding.assign_channels(len(chans), n, rank):
                out = d.process(chans[c])
                assert np.array_equal(out.view(np.uint32), base[c].view(np.uint32)), (rank, c)
    for c in (0, len(chans) - 1):
        check_close(base[c], oracle_best.run(chans[c], cfg), what=f"channel {c}")


@needs2
def test_context_on_another_device_than_the_current_one():
    """the C ABI launches, times and checks on the CONTEXT's device whatever the caller's current
    device is (fmd_time_device recorded its events on the wrong device)"""
    import torch
    torch.cuda.set_device(0)
    raw = synth.config_signal("cfg2", 4)
    with fm.Demodulator(device=0, sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0) as d0:
        want = d0.process(raw).copy()
    d_in = torch.from_numpy(raw).to("cuda:1")
    d_out = torch.empty(raw.size // 4, dtype=torch.float32, device="cuda:1")
    with fm.Demodulator(device=1, sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0) as d1:
        assert torch.cuda.current_device() == 0
        ms = d1.time_device(d_in.data_ptr(), raw.size, d_out.data_ptr(), 0, 3)
        assert ms > 0.0
        d1.check(0)
        assert torch.cuda.current_device() == 0
    got = d_out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@needs2
def test_quality_stream_split_over_gpus(scratch):
    """north_star: "a single long stream is split into overlapped chunks with warm-up state handoff":
    -Q over FMD_DEVICES=all must match the serial oracle like the one-GPU stream does."""
    raw = synth.config_signal("cfg3", 96)
    src = scratch / "in.u8"
    raw.tofile(src)
    cfg = po.Config(fs=48000.0, fc_out=15000.0)
    want = po.port().quality(cfg, 2.4e6).run(raw)
    for devs in ("0", "all"):
        dst = scratch / f"q_{devs}.f32"
        env = dict(os.environ, FMD_DEVICES=devs, FMD_SLOT_MB="2", FMD_STATS="1")
        r = subprocess.run([CLI, "-i", str(src), "-o", str(dst), "-Q", "-r2400000", "-S48000", "-l15000"],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        if devs == "all":
            assert f"{_ndev()} GPU(s)" in r.stderr, r.stderr
        out = np.fromfile(dst, np.float32)
        assert out.size == want.size
        err = float(np.abs(out - want).max())
        assert err <= 2e-5, (devs, err)
