"""The slice schedule of fmd_demod_host (demodulator_b200/csrc/fmd_slices.h), checked without a GPU: the
header is plain C, compiled here into a two-function shared object and driven through ctypes.  Whatever
the call size, chunk size and slot size: the slices cover the call exactly and in order, all but the last
are whole chunks, none exceeds a slot; graded, they start at 4 MiB (or one chunk), grow, then shrink."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MiB = 1 << 20


@pytest.fixture(scope="module")
def lib():
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    src, so = os.path.join(build, "slices_shim.c"), os.path.join(build, "libslices_shim.so")
    with open(src, "w") as f:
        f.write('#include "fmd_slices.h"\n'
                "size_t shim_first(size_t B, size_t slot) { return fmd_slice_first(B, slot); }\n"
                "size_t shim_len(unsigned n, size_t rem, size_t B, size_t slot, int ramp) { return fmd_slice_len(n, rem, B, slot, ramp); }\n")
    subprocess.run(["gcc", "-O1", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I",
                    os.path.join(ROOT, "demodulator_b200", "csrc"), "-o", so, src], check=True)
    L = C.CDLL(so)
    L.shim_first.restype = C.c_size_t
    L.shim_first.argtypes = [C.c_size_t, C.c_size_t]
    L.shim_len.restype = C.c_size_t
    L.shim_len.argtypes = [C.c_uint, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
    return L


def schedule(L, nbytes, B, slot, ramp):
    out, off, n = [], 0, 0
    while off < nbytes:
        ln = L.shim_len(n, nbytes - off, B, slot, ramp)
        assert 0 < ln <= nbytes - off
        out.append(ln)
        off += ln
        n += 1
        assert n < 100000
    return out


def host_slot(nbytes, B, slice_mb=64):
    """the slot size host_pipeline asks fmd_ring_create for (fmdemod.cu): the slice in whole chunks, the whole call if shorter"""
    s = slice_mb * MiB // B * B or B
    if s > nbytes:
        s = (nbytes + B - 1) // B * B or B
    return s


CASES = [(2 << 30, 262144), (16 * MiB, 262144), (64 * 262144 + 100, 262144), (262144, 262144), (12, 4), (7, 4),
         (3 * MiB + 5, 65536), (512 * MiB, 256 * MiB), (100 * MiB + 2, 8 * MiB), (65 * MiB, 262144), (4 * MiB + 262144, 262144)]


@pytest.mark.parametrize("nbytes,B", CASES)
@pytest.mark.parametrize("ramp", [0, 1])
def test_slices_cover_the_call_in_whole_chunks(lib, nbytes, B, ramp):
    slot = host_slot(nbytes, B)
    s = schedule(lib, nbytes, B, slot, ramp)
    assert sum(s) == nbytes
    assert all(x <= slot for x in s)
    assert all(x % B == 0 for x in s[:-1])  # only the call's last slice may end inside a chunk
    if not ramp:
        assert all(x == slot for x in s[:-1])


@pytest.mark.parametrize("nbytes,B", CASES)
def test_graded_slices_grow_then_shrink(lib, nbytes, B):
    slot = host_slot(nbytes, B)
    s = schedule(lib, nbytes, B, slot, 1)
    first = lib.shim_first(B, slot)
    assert first == min(max(4 * MiB // B * B, B), slot)
    assert s[0] == min(first, nbytes)
    peak = int(np.argmax(s))
    assert all(a <= b for a, b in zip(s[:peak], s[1:peak + 1]))      # up to the largest slice ...
    assert all(a >= b for a, b in zip(s[peak:-1], s[peak + 1:-1]))   # ... then down (the last one is the remainder)
    if len(s) > 2:
        assert s[-2] >= min(first, s[-2])  # no slice before the last falls below the smallest graded size
        assert all(x >= first for x in s[:-1])
    # grading costs a bounded number of extra slices: two geometric ramps
    equal = -(-nbytes // slot)
    assert len(s) <= equal + 2 * max(1, int(np.log2(max(slot // first, 1))) + 2)


def test_bench_sized_call(lib):
    """2 GiB of cfg2 input, the bench's e2e step: 4, 8, 16, 32 MiB, then 64 MiB slices, then halving"""
    B = 262144
    s = [x // MiB for x in schedule(lib, 2 << 30, B, 64 * MiB, 1)]
    assert s[:5] == [4, 8, 16, 32, 64]
    assert s[-1] <= 4 and s[-2] == 4
    assert max(s) == 64 and s.count(64) >= 25


def test_random_geometries(lib):
    rng = np.random.default_rng(11)
    for _ in range(300):
        B = int(rng.choice([4, 4096, 65536, 262144, 1 << 20, 8 << 20]))
        nbytes = int(rng.integers(1, 600 * MiB))
        slot = host_slot(nbytes, B, int(rng.choice([1, 16, 32, 64, 128])))
        for ramp in (0, 1):
            s = schedule(lib, nbytes, B, slot, ramp)
            assert sum(s) == nbytes and max(s) <= slot and all(x % B == 0 for x in s[:-1])
