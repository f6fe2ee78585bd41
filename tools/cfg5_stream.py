#!/usr/bin/env python
"""tools/cfg5_stream.py -- BASELINE configs[4]: ONE continuous IQ stream through stdin, chunked over
pinned double-buffered host I/O, here at full size (default 16 G samples = 32 GB in, 32 GB out).

    python tools/cfg5_stream.py [--gsamples 16] [--devices all] [--quality]

A producer thread writes the period-extended cfg3 signal (64 chunks = 16 MiB per period) into the
CLI's stdin; the CLI's stdout is read back period by period.  Checks ("parity on
a prefix + spot chunks, throughput on the full stream"): the first period against the oracle, and
EVERY later period bit-for-bit against the first (the input is periodic and no state crosses a
chunk, so the output must be) -- which is also the order check for the multi-GPU split.
With --quality the stream runs through -Q (state carried / warm-up segments): there the periods
are compared after the filters' transient, within 1e-5.
"""
import argparse
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demodulator_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (the checker of the first period; nothing here is timed against it)

ap = argparse.ArgumentParser()
ap.add_argument("--gsamples", type=float, default=16.0, help="stream length in 1e9 complex samples")
ap.add_argument("--devices", default="all")
ap.add_argument("--quality", action="store_true")
args = ap.parse_args()

PERIOD_CHUNKS = 64
base = synth.config_signal("cfg3", PERIOD_CHUNKS)          # 16 MiB, 8 Mi samples
n_periods = max(2, int(round(args.gsamples * 1e9 / (base.size // 2))))
total_samples = n_periods * (base.size // 2)
exe = os.environ.get("FMD_CLI") or os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")
# -Q: 2.4 MS/s -> 75 kS/s, M = 32 (a power of two, so that a signal period holds a whole number of outputs)
flags = ["-Q", "-r2400000", "-S75000", "-l15000"] if args.quality else ["-S48000", "-l15000"]
env = dict(os.environ, FMD_DEVICES=args.devices, FMD_STATS="1")
p = subprocess.Popen([exe, "-i", "-", "-o", "-"] + flags, stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                     stderr=subprocess.PIPE, env=env, bufsize=0)
try:
    import fcntl
    for f in (p.stdin, p.stdout):
        fcntl.fcntl(f.fileno(), 1031, 1 << 20)  # F_SETPIPE_SZ
except Exception:
    pass


def produce():
    mv = memoryview(base)
    for _ in range(n_periods):
        off = 0
        while off < len(mv):
            off += os.write(p.stdin.fileno(), mv[off:off + (4 << 20)])
    p.stdin.close()


t0 = time.time()
th = threading.Thread(target=produce)
th.start()
M = 32 if args.quality else 1
per_out = base.size // 2 // M if args.quality else base.size // 4      # outputs per period
per_bytes = per_out * 4
first = None
bad_periods, worst = 0, 0.0
buf = bytearray(per_bytes)
view = memoryview(buf)
got_periods = 0
while True:
    n = 0
    while n < per_bytes:
        r = p.stdout.readinto(view[n:])
        if not r:
            break
        n += r
    if n < per_bytes:
        tail = n
        break
    cur = np.frombuffer(buf, np.float32)
    if first is None:
        first = cur.copy()
    elif args.quality:
        if got_periods == 1:
            second = cur.copy()  # the first period holds the start-up transient of the stream
        else:
            worst = max(worst, float(np.abs(cur - second).max()))
    elif not np.array_equal(cur.view(np.uint32), first.view(np.uint32)):
        bad_periods += 1
    got_periods += 1
th.join()
p.wait()
dt = time.time() - t0
stats = p.stderr.read().decode().strip().splitlines()
if args.quality:
    want = po.port().quality(po.Config(fs=75000.0, fc_out=15000.0), 2.4e6).run(base)
    err0 = float(np.abs(first - want).max())
    ok = p.returncode == 0 and got_periods == n_periods and err0 <= 1e-4 and worst <= 1e-5
    print(f"cfg5 -Q stream: {total_samples / 1e9:.2f} G samples through stdin -> stdout on FMD_DEVICES={args.devices}: "
          f"{dt:.2f} s wall = {total_samples / dt / 1e6:.1f} MS/s; periods out {got_periods}/{n_periods}; first period vs serial "
          f"oracle {err0:.2e}; later periods vs the second, max-abs {worst:.2e}; {'OK' if ok else 'FAILED'}")
else:
    want = po.best().run(base, po.Config(fs=48000.0, fc_out=15000.0))
    err0 = float(np.abs(first - want).max())
    ok = p.returncode == 0 and got_periods == n_periods and bad_periods == 0 and err0 <= 1e-4
    print(f"cfg5 stream: {total_samples / 1e9:.2f} G samples ({2 * total_samples / 1e9:.1f} GB in, {2 * total_samples / 1e9:.1f} GB out) "
          f"through stdin -> stdout on FMD_DEVICES={args.devices}: {dt:.2f} s wall = {total_samples / dt / 1e6:.1f} MS/s; "
          f"periods out {got_periods}/{n_periods}, not bit-identical to the first: {bad_periods}; first period vs the "
          f"reference max-abs {err0:.2e}; {'OK' if ok else 'FAILED'}")
for ln in stats[-2:]:
    print("   ", ln)
sys.exit(0 if ok else 1)
