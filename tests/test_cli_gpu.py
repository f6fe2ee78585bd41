"""GPU: the outer boundary -- the C CLI (same flags and stream contract as the reference's main)
and the inner boundary -- the reference's OWN main linked against this backend's
processMatrix/allocateBuffer (oracle/_ref/demodulator_seam)."""
from __future__ import annotations

import os
import subprocess
import time

import numpy as np
import pytest

from demodulator_b200 import synth
from oracle import pyoracle as po
from helpers import B, check_close

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")
SEAM = os.path.join(ROOT, "oracle", "_ref", "demodulator_seam")



@pytest.fixture
def tmp_path():
    """A scratch directory WITHOUT '-' in its path: like the reference (src/main.c:128,136), the CLI
    takes any -i/-o argument that contains a '-' to mean stdin/stdout (pytest's own tmp_path has one)."""
    import shutil
    import tempfile
    from pathlib import Path
    d = tempfile.mkdtemp(prefix="fmd_", dir="/tmp")
    assert "-" not in d
    yield Path(d)
    shutil.rmtree(d, ignore_errors=True)


CASES = [
    ("cfg1", ["-S96000", "-l12500"], dict(fs=96000.0, fc_out=12500.0)),
    ("cfg2", ["-S96000", "-l6500", "-L12500"], dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
    ("cfg3", ["-S48000", "-l15000"], dict(fs=48000.0, fc_out=15000.0)),
    ("cfg2", ["-S96000", "-l12500", "-d5", "-L12500", "-m3"],
     dict(fs=96000.0, fc_out=12500.0, fc_in=12500.0, deg_out=5, mode=3)),
    ("cfg1", [], dict()),
    ("cfg2", ["-S96000", "-l6500", "-L12500", "-q3"], dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, mode=12)),
    ("cfg1", ["-S96000", "-l12500", "-q2"], dict(fs=96000.0, fc_out=12500.0, mode=8)),
    ("cfg1", ["-S96000", "-l12500", "-q1"], dict(fs=96000.0, fc_out=12500.0, mode=4)),
    ("cfg2", ["-S96000", "-l6500", "-L12500", "-q1"], dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, mode=4)),
]


@pytest.mark.parametrize("sig,flags,kw", CASES)
def test_cli_file_to_file(tmp_path, sig, flags, kw, oracle_best):
    raw = synth.config_signal(sig, 5)
    src, dst = tmp_path / "in.u8", tmp_path / "out.f32"
    raw.tofile(src)
    r = subprocess.run([CLI, "-i", str(src), "-o", str(dst)] + flags, capture_output=True, text=True,
                       env=dict(os.environ, FMD_SLOT_MB="1"))  # 1 MiB slots: several ring turns
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst, np.float32)
    assert out.size == raw.size // 4  # exactly bufSize/4 floats per bufSize bytes, no extra chunk
    check_close(out, oracle_best.run(raw, po.Config(**kw)), what=f"CLI {flags}")


def test_cli_stdin_stdout_pipe_and_tail(oracle_port):
    """-i - -o -: stdin/stdout; a slow producer (bytes trickle in) and a ragged tail."""
    raw = synth.config_signal("cfg3", 3)[: 2 * B + 1000 * 4 + 2]
    p = subprocess.Popen([CLI, "-i", "-", "-o", "-", "-S48000", "-l15000"], stdin=subprocess.PIPE,
                         stdout=subprocess.PIPE)
    import threading
    got: list[bytes] = []
    t = threading.Thread(target=lambda: got.append(p.stdout.read()))
    t.start()
    data = raw.tobytes()
    for off in range(0, len(data), 100000):
        p.stdin.write(data[off:off + 100000])
        p.stdin.flush()
        time.sleep(0.01)
    p.stdin.close()
    t.join()
    assert p.wait() == 0
    out = np.frombuffer(got[0], np.float32)
    cfg = po.Config(fs=48000.0, fc_out=15000.0)
    assert out.size == (raw.size & ~3) // 4
    check_close(out[: 2 * B // 4], oracle_port.run(raw[: 2 * B], cfg), what="whole chunks")
    check_close(out[2 * B // 4:], oracle_port.run(raw[2 * B: 2 * B + 4000], cfg, chunk_bytes=4000), what="tail")


def test_cli_dash_in_name_means_stdio(tmp_path, oracle_port):
    """src/main.c:128,136: strstr(optarg, "-") -- a file NAME containing '-' selects stdin/stdout."""
    raw = synth.config_signal("cfg1", 1)
    r = subprocess.run([CLI, "-i", "my-input.u8", "-o", "my-output.f32", "-S96000"], input=raw.tobytes(),
                       capture_output=True)
    assert r.returncode == 0 and not os.path.exists("my-output.f32")
    out = np.frombuffer(r.stdout, np.float32)
    check_close(out, oracle_port.run(raw, po.Config(fs=96000.0)), what="stdio")


def test_cli_bufsize_flag(tmp_path, oracle_port):
    raw = synth.config_signal("cfg1", 2)
    src, dst = tmp_path / "in.u8", tmp_path / "out.f32"
    raw.tofile(src)
    assert subprocess.run([CLI, "-i", str(src), "-o", str(dst), "-S96000", "-b-4"]).returncode == 0
    out = np.fromfile(dst, np.float32)
    check_close(out, oracle_port.run(raw, po.Config(fs=96000.0), chunk_bytes=B >> 4), what="-b-4")


def test_cli_rejects_undefined_configs(tmp_path):
    src = tmp_path / "in.u8"
    synth.config_signal("cfg1", 1).tofile(src)
    r = subprocess.run([CLI, "-i", str(src), "-o", str(tmp_path / "o"), "-D3"], capture_output=True, text=True)
    assert r.returncode != 0 and "-D without -L" in r.stderr
    r = subprocess.run([CLI, "-i", str(src), "-o", str(tmp_path / "o"), "-q2", "-b-6"], capture_output=True, text=True)
    assert r.returncode != 0 and "-q2" in r.stderr


@pytest.mark.skipif(not os.path.exists(SEAM), reason="seam binary (reference main + this backend) not prebuilt")
@pytest.mark.parametrize("flags,kw", [(["-S96000", "-l12500"], dict(fs=96000.0, fc_out=12500.0)),
                                      (["-S96000", "-l6500", "-L12500"],
                                       dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0))])
def test_reference_main_over_this_backend(tmp_path, flags, kw, oracle_best):
    """The reference's unmodified src/main.c (reader loop, semaphores, getopt) driving THIS
    backend through processMatrix/allocateBuffer: drop-in at the seam the reference itself uses
    to swap its CPU and CUDA backends (src/main.c:44-47)."""
    raw = synth.config_signal("cfg2", 6)
    src, dst = tmp_path / "in.u8", tmp_path / "out.f32"
    raw.tofile(src)
    r = subprocess.run([SEAM, "-i", str(src), "-o", str(dst)] + flags, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst, np.float32)
    assert out.size == raw.size // 4
    check_close(out, oracle_best.run(raw, po.Config(**kw)), what=f"seam {flags}")


def test_cli_all_gpus_round_robin(tmp_path, oracle_best):
    """FMD_DEVICES=all: the slots of ONE stream go round-robin over every visible GPU and come
    back in order (chunk-aligned ranges need no overlap and no hand-off state)."""
    raw = synth.config_signal("cfg2", 24)
    src, dst = tmp_path / "in.u8", tmp_path / "out.f32"
    raw.tofile(src)
    r = subprocess.run([CLI, "-i", str(src), "-o", str(dst), "-S96000", "-l6500", "-L12500"], capture_output=True,
                       text=True, env=dict(os.environ, FMD_DEVICES="all", FMD_SLOT_MB="1", FMD_STATS="1"))
    assert r.returncode == 0, r.stderr
    out = np.fromfile(dst, np.float32)
    check_close(out, oracle_best.run(raw, po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0)), what="all GPUs")


def test_rtltcp_feed_into_the_cli(oracle_best):
    """acquisition adapter end to end: stand-in rtl_tcp server -> rtltcp_feed -> demodulator -i - (GPU) ==
    the reference on the same bytes"""
    from test_rtltcp_feed import FEED, FakeRtlTcp
    raw = synth.config_signal("cfg3", 6)
    srv = FakeRtlTcp(raw.tobytes(), pause_after=3 * B)
    srv.start()
    r = subprocess.run(f"{FEED} -p {srv.port} -s 2400000 -f 100e6 | {CLI} -i - -o - -S48000 -l15000", shell=True,
                       capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    out = np.frombuffer(r.stdout, np.float32)
    check_close(out, oracle_best.run(raw, po.Config(fs=48000.0, fc_out=15000.0)), what="rtl_tcp -> feed -> CLI")
