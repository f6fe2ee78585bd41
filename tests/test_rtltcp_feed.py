"""The acquisition adapter (demodulator_b200/csrc/rtltcp_feed.c) against a stand-in
rtl_tcp server: greeting consumed, the opening commands of the reference's helper sent
(/root/reference/rtltcp.py:71-78: gain mode, AGC mode, bias tee, sample rate) plus the ones asked for on
the command line and through the control channel, IQ bytes forwarded unchanged -- and, with the
stand-in backend of test_cli_host_logic.py, all the way through the C host as a live stream."""
import os
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FEED = os.path.join(ROOT, "demodulator_b200", "bin", "rtltcp_feed")


class FakeRtlTcp(threading.Thread):
    """greets like rtl_tcp, records 5-byte commands, streams `payload` after the first command arrived"""

    def __init__(self, payload: bytes, pause_after: int = 0, throttle: float = 0.0):
        super().__init__(daemon=True)
        self.throttle = throttle
        self.srv = socket.socket()
        self.srv.bind(("127.0.0.1", 0))
        self.srv.listen(1)
        self.port = self.srv.getsockname()[1]
        self.payload, self.pause_after = payload, pause_after
        self.commands = []
        self.lock = threading.Lock()

    def run(self):
        conn, _ = self.srv.accept()
        conn.sendall(b"RTL0" + struct.pack(">II", 5, 29))
        stop = threading.Event()

        def listen():
            buf = b""
            while not stop.is_set():
                try:
                    d = conn.recv(4096)
                except OSError:
                    break
                if not d:
                    break
                buf += d
                while len(buf) >= 5:
                    with self.lock:
                        self.commands.append(struct.unpack(">BI", buf[:5]))
                    buf = buf[5:]
        t = threading.Thread(target=listen, daemon=True)
        t.start()
        for _ in range(200):  # wait for the opening commands
            with self.lock:
                if len(self.commands) >= 4:
                    break
            time.sleep(0.01)
        off = 0
        while off < len(self.payload):
            n = min(200000, len(self.payload) - off)
            try:
                conn.sendall(self.payload[off:off + n])
            except OSError:  # the client went away (the "q" line of the control channel)
                break
            off += n
            if self.throttle:
                time.sleep(self.throttle)
            if self.pause_after and off >= self.pause_after:
                time.sleep(0.3)
                self.pause_after = 0
        time.sleep(0.1)
        stop.set()
        try:
            conn.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        conn.close()


@pytest.fixture(scope="module")
def feed():
    if not os.path.exists(FEED):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "demodulator_b200", "csrc"),
                        os.path.join(ROOT, "demodulator_b200", "bin", "rtltcp_feed")], check=True)
    return FEED


def test_feed_tunes_and_forwards(feed, tmp_path):
    payload = np.random.default_rng(1).integers(0, 256, 3_000_001, dtype=np.uint8).tobytes()
    srv = FakeRtlTcp(payload, pause_after=1_000_000)
    srv.start()
    ctl = tmp_path / "ctl"
    ctl.write_text("")
    p = subprocess.Popen([feed, "-p", str(srv.port), "-f", "100.1e6", "-s", "2400000", "-g", "280", "-P", "-3", "-c", str(ctl)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    time.sleep(0.15)
    with open(ctl, "a") as f:  # a retune while the stream runs, by name and by number
        f.write("SET_FREQUENCY 99500000\n0x0d 7\n")
    out, err = p.communicate(timeout=30)
    srv.join(timeout=5)
    assert p.returncode == 0, err.decode()
    assert out == payload
    cmds = srv.commands
    assert cmds[:4] == [(0x03, 1), (0x08, 0), (0x0e, 0), (0x02, 2400000)]
    assert (0x01, 100100000) in cmds and (0x04, 280) in cmds and (0x05, 0xFFFFFFFD) in cmds
    assert (0x01, 99500000) in cmds and (0x0d, 7) in cmds
    assert b"tuner type 5, 29 gain steps" in err


def test_feed_into_the_c_host_as_a_live_stream(feed, tmp_path):
    """rtltcp_feed | demodulator -i - : a socket-fed pipe is a stream to the host (slots cut short when the
    producer pauses); run against the stand-in backend, whose outputs name the input position"""
    stub = os.path.join(ROOT, "tests", "_build", "demodulator_stub")
    if not os.path.exists(stub):
        subprocess.run(["gcc", "-O1", "-std=gnu99", "-I", os.path.join(ROOT, "include"), "-o", stub,
                        os.path.join(ROOT, "demodulator_b200", "csrc", "demodulator.c"),
                        os.path.join(ROOT, "tests", "stub", "stub_backend.c"), "-lpthread", "-lm"], check=True)
    nbytes = 4 * 262144 + 262144 // 2
    words = np.arange(nbytes // 4, dtype=np.uint32)
    srv = FakeRtlTcp(words.tobytes(), pause_after=2 * 262144)
    srv.start()
    env = dict(os.environ, FMD_SLOT_MB="1", FMD_SLOTS="2")
    r = subprocess.run(f"{feed} -p {srv.port} -s 2400000 | {stub} -i - -o - -S48000 -l15000", shell=True, capture_output=True,
                       env=env, timeout=60)
    assert r.returncode == 0, r.stderr.decode()
    got = np.frombuffer(r.stdout, np.float32)
    assert np.array_equal(got, (words & 0xffffff).astype(np.float32))


def test_control_channel_extended_names_and_quit(feed, tmp_path):
    """the command names of the rtl_tcp forks that the reference's helper lists (rtltcp.py:32-63) and its
    'q' line: the feed ends the stream itself, exit status 0, what it forwarded is a prefix of the stream"""
    payload = np.random.default_rng(2).integers(0, 256, 40_000_000, dtype=np.uint8).tobytes()
    srv = FakeRtlTcp(payload, throttle=0.01)  # ~20 MB/s: several seconds of stream
    srv.start()
    ctl = tmp_path / "ctl"
    ctl.write_text("")
    p = subprocess.Popen([feed, "-p", str(srv.port), "-A", "-T", "-c", str(ctl)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    time.sleep(0.3)
    with open(ctl, "a") as f:
        f.write("SET_FREQ_HI32 1\nset_sideband 1\nbogus 3\n")
    time.sleep(0.4)  # let the commands reach the server before the feed closes the socket under its stream
    with open(ctl, "a") as f:
        f.write("q\n")
    out, err = p.communicate(timeout=30)
    srv.join(timeout=10)
    assert p.returncode == 0, err.decode()
    assert 0 < len(out) < len(payload) and payload.startswith(out)
    cmds = srv.commands
    assert cmds[:4] == [(0x03, 0), (0x08, 1), (0x0e, 1), (0x02, 250000)]  # -A: automatic gain, AGC on; -T: bias tee
    assert (0x56, 1) in cmds and (0x47, 1) in cmds
    assert b"not a command: bogus 3" in err
