"""Host logic of the C CLI (demodulator_b200/csrc/demodulator.c) without a GPU: the slot ring,
the reader / writer threads and the positioned I/O, linked against a stand-in backend
(tests/stub/stub_backend.c) whose outputs identify the input position they came from.

What is pinned here: slot numbers n and n + nslots must never be confused whatever
the number of I/O threads and ring slots; an inherited stdout / stdin that does not stand at byte 0,
or is in append mode, must behave like the reference's sequential fread / fwrite
(/root/reference/src/main.c:72-91, src/matrix.c:256,268)."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "demodulator_b200", "csrc", "demodulator.c")
STUB = os.path.join(ROOT, "tests", "stub", "stub_backend.c")
BUILD = os.path.join(ROOT, "tests", "_build")
EXE = os.path.join(BUILD, "demodulator_stub")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(BUILD, exist_ok=True)
    subprocess.run(["gcc", "-O1", "-g", "-std=gnu99", "-Wall", "-I", os.path.join(ROOT, "include"),
                    "-o", EXE, SRC, STUB, "-lpthread", "-lm"], check=True)
    return EXE


@pytest.fixture()
def tmp_path():
    """a scratch directory whose name has no '-': any -i/-o argument containing '-' means
    stdin/stdout (src/main.c:128,136), and pytest's own tmp_path is full of them"""
    d = tempfile.mkdtemp(prefix="fmdhost", dir=tempfile.gettempdir().replace("-", "") if "-" in tempfile.gettempdir() else None)
    assert "-" not in d
    from pathlib import Path
    yield Path(d)
    shutil.rmtree(d, ignore_errors=True)


def make_input(path, nbytes):
    """little-endian counters: word k holds k, so output float k must be k & 0xffffff"""
    words = np.arange(nbytes // 4, dtype=np.uint32)
    raw = words.view(np.uint8)
    tail = np.arange(nbytes - raw.size, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(raw.tobytes())
        f.write(tail.tobytes())
    return (words & 0xffffff).astype(np.float32)


def run(exe, args, env_extra, **kw):
    env = dict(os.environ, STUB_JITTER="1", **env_extra)
    return subprocess.run([exe] + args, env=env, timeout=120, **kw)


@pytest.mark.parametrize("slots,threads,devices,mmap_out", [(2, 8, 1, True), (3, 4, 1, False), (2, 3, 2, True), (5, 7, 3, False),
                                                            (3, 16, 4, True), (2, 1, 1, True), (2, 8, 1, False)])
def test_file_to_file_keeps_order(exe, tmp_path, slots, threads, devices, mmap_out):
    # 1 MiB slots, > 4 * nslots slots of data, ragged tail (not a multiple of the chunk or of 4)
    nbytes = 41 * (1 << 20) + 262144 * 2 + 1234 + 3
    src, dst = str(tmp_path / "in.u8"), str(tmp_path / "out.f32")
    want = make_input(src, nbytes)
    env = dict(FMD_SLOT_MB="1", FMD_SLOTS=str(slots), FMD_IO_THREADS=str(threads),
               STUB_DEVICES=str(devices), FMD_DEVICES="all", FMD_NO_MMAP="0" if mmap_out else "1")
    for _ in range(3):
        r = run(exe, ["-i", src, "-o", dst], env)
        assert r.returncode == 0
        got = np.fromfile(dst, dtype=np.float32)
        assert got.size == nbytes // 4
        bad = np.nonzero(got != want[: got.size])[0]
        assert bad.size == 0, f"first misplaced output at float {bad[0]} (slot {bad[0] * 4 >> 20})"


def test_stream_in_stream_out_keeps_order(exe, tmp_path):
    nbytes = 9 * (1 << 20) + 262144
    src = str(tmp_path / "in.u8")
    want = make_input(src, nbytes)
    env = dict(os.environ, STUB_JITTER="1", FMD_SLOT_MB="1", FMD_SLOTS="2", STUB_DEVICES="2", FMD_DEVICES="all")
    r = subprocess.run(f"cat {src} | {exe} -i - -o -", shell=True, env=env, capture_output=True, timeout=120)
    assert r.returncode == 0
    got = np.frombuffer(r.stdout, dtype=np.float32)
    assert np.array_equal(got, want[: got.size]) and got.size == nbytes // 4


def test_inherited_stdout_keeps_what_was_written_before(exe, tmp_path):
    nbytes = 5 * (1 << 20)
    src, dst = str(tmp_path / "in.u8"), str(tmp_path / "out.bin")
    want = make_input(src, nbytes)
    env = dict(os.environ, FMD_SLOT_MB="1", FMD_SLOTS="2", FMD_IO_THREADS="4")
    # header before, trailer after: both must survive, output in between
    r = subprocess.run(f"{{ printf 'HEADER--'; {exe} -i {src} -o -; printf 'TRAILER-'; }} > {dst}", shell=True, env=env,
                       timeout=120)
    assert r.returncode == 0
    blob = open(dst, "rb").read()
    assert blob[:8] == b"HEADER--" and blob[-8:] == b"TRAILER-"
    assert np.array_equal(np.frombuffer(blob[8:-8], dtype=np.float32), want)


def test_append_mode_stdout_is_written_in_order(exe, tmp_path):
    nbytes = 6 * (1 << 20)
    src, dst = str(tmp_path / "in.u8"), str(tmp_path / "out.bin")
    want = make_input(src, nbytes)
    open(dst, "wb").write(b"OLD-DATA")
    env = dict(os.environ, STUB_JITTER="1", FMD_SLOT_MB="1", FMD_SLOTS="2", FMD_IO_THREADS="4")
    r = subprocess.run(f"{exe} -i {src} -o - >> {dst}", shell=True, env=env, timeout=120)
    assert r.returncode == 0
    blob = open(dst, "rb").read()
    assert blob[:8] == b"OLD-DATA"
    assert np.array_equal(np.frombuffer(blob[8:], dtype=np.float32), want)


def test_inherited_stdin_is_read_from_where_it_stands(exe, tmp_path):
    nbytes = 4 * (1 << 20)
    src, dst = str(tmp_path / "in.u8"), str(tmp_path / "out.f32")
    want = make_input(src, nbytes)
    skip = 1 << 20
    env = dict(os.environ, FMD_SLOT_MB="1", FMD_SLOTS="2", FMD_IO_THREADS="4")
    # the first MiB is consumed by head; the demodulator must start after it and leave the
    # descriptor at the end of what it read
    r = subprocess.run(f"{{ head -c {skip} > /dev/null; {exe} -i - -o {dst}; wc -c; }} < {src}", shell=True, env=env,
                       capture_output=True, timeout=120)
    assert r.returncode == 0
    got = np.fromfile(dst, dtype=np.float32)
    assert np.array_equal(got, want[skip // 4:])
    assert int(r.stdout.split()[0]) == 0  # nothing left for the next reader of the descriptor


@pytest.mark.parametrize("from_file", [True, False])
def test_segmented_quality_stream_hands_every_gpu_its_warm_up_prefix(exe, tmp_path, from_file):
    """-Q over several GPUs: every slot is a segment that starts `fmd_warmup_samples` early.  The
    stand-in backend checks that the prefix it is handed is exactly the stream before the payload
    (and fails the submit otherwise); the outputs must still come back in order."""
    nbytes = 23 * (1 << 20) + 262144
    src, dst = str(tmp_path / "in.u8"), str(tmp_path / "out.f32")
    want = make_input(src, nbytes)
    env = dict(os.environ, STUB_JITTER="1", FMD_SLOT_MB="1", FMD_SLOTS="2", FMD_IO_THREADS="4", STUB_DEVICES="3",
               FMD_DEVICES="all", STUB_WARM="5000")
    cmd = f"{exe} -i {src} -o {dst} -Q -r2400000 -S48000" if from_file else f"cat {src} | {exe} -i - -o {dst} -Q -r2400000 -S48000"
    r = subprocess.run(cmd, shell=True, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(dst, dtype=np.float32)
    assert got.size == nbytes // 4 and np.array_equal(got, want)


def test_segmented_stream_with_a_stalling_producer(exe, tmp_path):
    """a live source: the producer pauses, slots are cut short at chunk boundaries; every segment's
    warm-up prefix must still be exactly the bytes that precede it (the stand-in backend checks)"""
    import threading
    import time
    nbytes = 6 * (1 << 20) + 3 * 262144
    src = str(tmp_path / "in.u8")
    want = make_input(src, nbytes)
    data = open(src, "rb").read()
    env = dict(os.environ, FMD_SLOT_MB="1", FMD_SLOTS="2", STUB_DEVICES="2", FMD_DEVICES="all", STUB_WARM="3000")
    p = subprocess.Popen([exe, "-i", "-", "-o", "-", "-Q", "-r2400000", "-S48000"], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, env=env)
    got = []
    rd = threading.Thread(target=lambda: got.append(p.stdout.read()))
    rd.start()
    off = 0
    for piece in (262144, 3 * 262144 + 100, 700000, 262144 * 5 - 100, nbytes):
        end = min(nbytes, off + piece)
        p.stdin.write(data[off:end])
        p.stdin.flush()
        off = end
        time.sleep(0.05)
        if off >= nbytes:
            break
    p.stdin.close()
    rd.join()
    assert p.wait(timeout=60) == 0, p.stderr.read().decode()
    out = np.frombuffer(got[0], dtype=np.float32)
    assert out.size == nbytes // 4 and np.array_equal(out, want)
