"""The measured-and-not-shipped kernel variants stay buildable (they are build options of the product
source), and the TMA one really is a TMA ring: its SASS holds the bulk
copy and the mbarrier instructions (B200_PROFILING.md: `cp.async.bulk` shows up as UBLKCP)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "demodulator_b200", "csrc")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _compile(tmp_path, *defines):
    obj = str(tmp_path / "inst.o")
    cmd = [NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
           "-I", CSRC, "-DFMD_INST_GROUP=9", *defines, "-c", "-o", obj, os.path.join(CSRC, "fmd_fast_inst.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout


@pytest.mark.skipif(not os.path.exists(NVCC), reason="no nvcc")
def test_tma_ring_variant_builds_and_uses_bulk_copies(tmp_path):
    sass = _compile(tmp_path, "-DFMD_TMA=1")
    assert "UBLKCP" in sass and "SYNCS" in sass and "ELECT" in sass
    assert "LDGSTS" not in sass  # no per-lane cp.async left in that build


@pytest.mark.skipif(not os.path.exists(NVCC), reason="no nvcc")
def test_product_build_uses_the_cp_async_ring_and_packed_f32x2(tmp_path):
    sass = _compile(tmp_path, "-DFMD_EARLY_LDS=1")  # also keeps that option compiling
    assert "LDGSTS" in sass and "FFMA2" in sass and "UBLKCP" not in sass
