#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (authoring container only).

Needs /root/reference (compiled in place by `make -C oracle ref`).  For every case
it runs

  (1) the reference's functions through oracle/_ref/libdemod_ref.so (ref_shim.c),
  (2) the reference's CLI binary oracle/_ref/demodulator_ref, black-box, fed through
      a pipe whose writer stays open after the last byte (the reference's EOF
      handling races on args->bufSize -- SURVEY.md A.6 -- so only the first K
      whole chunks of its output are compared and the trailing stale chunk is
      dropped),

checks (1) == (2) bit for bit, and stores the input recipe, sha256 digests and a
strided subset of the output as the fixture.  The fixtures are what pins the
oracle port and the CUDA path on the GPU box, where /root/reference is absent.

Usage:  python tests/golden/make_golden.py [case ...]     (no argument: every fixture)
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from demodulator_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
B = synth.CHUNK_BYTES

# name, signal config, chunks, CLI flags, Config kwargs
CASES = [
    ("cfg1", "cfg1", 3, ["-S96000", "-l12500"], dict(fs=96000.0, fc_out=12500.0)),
    ("cfg2", "cfg2", 3, ["-S96000", "-l6500", "-L12500"], dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
    ("cfg3", "cfg3", 3, ["-S48000", "-l15000"], dict(fs=48000.0, fc_out=15000.0)),
    ("default", "cfg1", 2, [], dict()),
    ("d5_L", "cfg2", 2, ["-S96000", "-l12500", "-d5", "-L12500"],
     dict(fs=96000.0, fc_out=12500.0, fc_in=12500.0, deg_out=5)),
    ("m3_L", "cfg2", 2, ["-S96000", "-l9600", "-m3", "-L12500"],
     dict(fs=96000.0, fc_out=9600.0, fc_in=12500.0, mode=3)),
    ("m1_d4", "cfg1", 2, ["-S96000", "-l12500", "-m1", "-d4", "-e2"],
     dict(fs=96000.0, fc_out=12500.0, deg_out=4, mode=1, eps=0.2)),
    # eps >= log10(2): the reference's acosh argument drops below 1 and every coefficient is NaN
    ("m1_e5_nan", "cfg1", 1, ["-S96000", "-l12500", "-m1", "-e5"],
     dict(fs=96000.0, fc_out=12500.0, mode=1, eps=0.5)),
    ("D7_d4_L", "cfg2", 2, ["-S96000", "-l9600", "-d4", "-D7", "-L12500", "-m2"],
     dict(fs=96000.0, fc_out=9600.0, fc_in=12500.0, deg_out=4, deg_in=7, mode=2)),
    # -q3: normalizeInput instead of shiftOrigin (mode bits 2-3)
    ("q3", "cfg1", 2, ["-S96000", "-l12500", "-q3"], dict(fs=96000.0, fc_out=12500.0, mode=12)),
    ("q3_L", "cfg2", 2, ["-S96000", "-l6500", "-L12500", "-q3"],
     dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, mode=12)),
    # -q2: highpassDc (shiftOrigin + a Butterworth highpass at 1 Hz through applyComplexFilter); its
    # overrun lands in the accumulator of the next stage.  All -q2 cases share -S and the chunk size:
    # the reference designs that filter and allocates its scratch once per process.
    ("q2", "cfg1", 2, ["-S96000", "-l12500", "-q2"], dict(fs=96000.0, fc_out=12500.0, mode=8)),
    ("q2_L", "cfg2", 2, ["-S96000", "-l6500", "-L12500", "-q2"],
     dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, mode=8)),
    # -q1: correctIq keeps its offsets in a function-local static, so the shim cannot be driven case
    # by case in one process: these outputs are the reference CLI's alone, and the oracle PORT is what
    # is checked against them here.
    ("q1", "cfg1", 3, ["-S96000", "-l12500", "-q1"], dict(fs=96000.0, fc_out=12500.0, mode=4)),
    ("q1_L", "cfg2", 3, ["-S96000", "-l6500", "-L12500", "-q1"],
     dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, mode=4)),
    ("q1_cfg3", "cfg3", 2, ["-S48000", "-l15000", "-q1"], dict(fs=48000.0, fc_out=15000.0, mode=4)),
    ("q2_d5_D4_L_m3", "cfg2", 2, ["-S96000", "-l9600", "-L12500", "-q2", "-d5", "-D4", "-m3"],
     dict(fs=96000.0, fc_out=9600.0, fc_in=12500.0, deg_out=5, deg_in=4, mode=8 | 3)),
]

# byte patterns exercising the int8 wrap (255 -> -128), signed zeros and the +-pi cut
def edge_bytes() -> np.ndarray:
    rng = np.random.default_rng(77)
    raw = rng.integers(0, 256, 2 * B, dtype=np.uint8)
    pat = np.array([
        130, 127, 124, 127,   # (3,0),(-3,0)  -> -pi
        124, 127, 130, 127,   # (-3,0),(3,0)  -> +pi
        127, 127, 122, 122,   # (0,0),(-5,-5) -> +pi
        127, 127, 127, 127,   # (0,0),(0,0)   -> 0
        255, 255, 0, 0,       # wrap: 255 -> -128, 0 -> -127
        255, 0, 255, 0, 0, 255, 255, 127, 127, 255, 128, 126,
    ], np.uint8)
    raw[4096:4096 + pat.size] = pat
    raw[B - pat.size:B] = pat          # at a chunk end
    raw[B:B + pat.size] = pat          # at a chunk start
    raw[B + 65536:B + 65536 + 4096] = 127  # a run of exact zeros
    raw[B + 131072:B + 131072 + 4096] = 255
    return raw


def run_binary(raw: np.ndarray, flags: list[str]) -> np.ndarray:
    """Black-box run of the reference CLI through a pipe with delayed close."""
    p = subprocess.Popen([po.REF_BIN, "-i", "-", "-o", "-"] + flags,
                         stdin=subprocess.PIPE, stdout=subprocess.PIPE)
    import threading
    chunks: list[bytes] = []
    t = threading.Thread(target=lambda: chunks.append(p.stdout.read()))
    t.start()
    p.stdin.write(raw.tobytes())
    p.stdin.flush()
    time.sleep(1.0)
    p.stdin.close()
    t.join()
    p.wait()
    return np.frombuffer(chunks[0], np.float32)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    po.build(ref=True)
    R = po.ref()
    only = set(sys.argv[1:])
    for name, sig_name, nchunks, flags, kw in CASES:
        if only and name not in only:
            continue
        raw = synth.config_signal(sig_name, nchunks)
        cfg = po.Config(**kw)
        out = (po.port() if ((cfg.mode >> 2) & 3) == 1 else R).run(raw, cfg)
        outb = run_binary(raw, flags)
        nb = min(outb.size, out.size)
        assert nb >= out.size, (name, outb.size, out.size)
        same = np.array_equal(out.view(np.uint32), outb[:out.size].view(np.uint32))
        assert same, f"{name}: CLI binary and shim disagree"
        _save(name, raw, out, cfg, flags, sig=dict(kind=sig_name, nchunks=nchunks), binary_checked=True)
        print(f"{name}: {out.size} floats, binary == shim, extra floats from CLI = {outb.size - out.size}")
    if only:
        return
    raw = edge_bytes()
    for name, flags, kw in (("edge", ["-S96000", "-l12500"], dict(fs=96000.0, fc_out=12500.0)),
                            ("edge_L", ["-S96000", "-l6500", "-L12500"],
                             dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0))):
        cfg = po.Config(**kw)
        out = R.run(raw, cfg)
        outb = run_binary(raw, flags)
        assert np.array_equal(out.view(np.uint32), outb[:out.size].view(np.uint32)), name
        _save(name, raw, out, cfg, flags, sig=dict(kind="edge_bytes", nchunks=2), binary_checked=True,
              keep_raw=True)
        print(f"{name}: ok")
    # filter coefficient table (SURVEY.md A.2 and more)
    rows = []
    for kind in range(4):
        for deg in range(1, 9):
            for fc, fs in ((12500., 96000.), (6500., 96000.), (15000., 48000.), (12500., 125000.),
                           (9600., 96000.), (1., 96000.)):
                for eps in (0.3, 0.2):
                    if kind in (0, 2) and eps != 0.3:
                        continue
                    sos = R.design(kind, deg, fc, fs, eps)
                    pad = np.zeros((4, 6), np.float32)
                    pad[:sos.shape[0]] = sos
                    rows.append((kind, deg, fc, fs, eps, pad))
    np.savez_compressed(os.path.join(HERE, "design.npz"),
                        kind=np.array([r[0] for r in rows], np.int32),
                        degree=np.array([r[1] for r in rows], np.int32),
                        fc=np.array([r[2] for r in rows]), fs=np.array([r[3] for r in rows]),
                        eps=np.array([r[4] for r in rows]),
                        sos=np.stack([r[5] for r in rows]))
    print(f"design.npz: {len(rows)} filters")


def _save(name, raw, out, cfg, flags, sig, binary_checked, keep_raw=False) -> None:
    n_out = B // 4
    nchunks = out.size // n_out
    o = out.reshape(nchunks, n_out)
    extra = {}
    if keep_raw:
        extra["raw"] = raw
        extra["out_full"] = out
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"),
        flags=np.array(" ".join(flags)),
        cfg=np.array([cfg.fs, cfg.fc_in, cfg.fc_out, cfg.eps, cfg.deg_in, cfg.deg_out, cfg.mode], np.float64),
        sig_kind=np.array(sig["kind"]), nchunks=np.array(nchunks),
        raw_sha256=np.array(sha(raw)), out_sha256=np.array(sha(out)),
        head=o[:, :64].copy(), tail=o[:, -64:].copy(), strided=out[::97].copy(),
        binary_checked=np.array(binary_checked), **extra)


if __name__ == "__main__":
    main()
