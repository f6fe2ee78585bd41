"""CPU: the C-ABI library loads, exports what include/fmdemod.h declares, and refuses to compute
without a GPU (no CPU fallback anywhere in the product)."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import pytest

import demodulator_b200 as fm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "demodulator_b200", "lib")
CLI = os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")


def _have_gpu() -> bool:
    try:
        return fm.device_count() > 0
    except Exception:
        return False


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "fmdemod.h")).read()
    names = sorted(set(re.findall(r"\b(fmd_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(fm.lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_compat_seam_exports():
    lib = ctypes.CDLL(os.path.join(LIB, "libfmd_compat.so"))
    assert hasattr(lib, "processMatrix") and hasattr(lib, "allocateBuffer")


def test_no_cpu_fallback_in_product_sources():
    """The product may not reach into oracle/ (a CPU path there would void every parity claim)."""
    bad = []
    for base in ("demodulator_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle|fm_oracle\.h|pyoracle|libfm_oracle|libdemod_ref", txt, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_defaults_match_reference_cli():
    from demodulator_b200.api import _Config
    c = _Config()
    fm.load_library().fmd_config_default(ctypes.byref(c))
    # src/main.c:102-113
    assert (c.sample_rate, c.lowpass_in, c.lowpass_out, c.epsilon) == (125000.0, 0.0, 12500.0, 0.3)
    assert (c.in_degree, c.out_degree, c.mode, c.chunk_bytes) == (0, 3, 0x10, 262144)


@pytest.mark.skipif(_have_gpu(), reason="a GPU is present")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(fm.FmdError, match="no CUDA device|no CPU path"):
        fm.Demodulator()


def test_design_rejects_bad_arguments():
    with pytest.raises(fm.FmdError):
        fm.design_sos(0, 0, 12500., 96000.)
    with pytest.raises(fm.FmdError):
        fm.design_sos(0, 17, 12500., 96000.)
    with pytest.raises(fm.FmdError):
        fm.design_sos(7, 3, 12500., 96000.)


def test_cli_argc_contract():
    """src/main.c:122-123: fewer than 3 arguments -> return -1 (exit status 255)."""
    assert os.path.exists(CLI)
    assert subprocess.run([CLI]).returncode == 255
    assert subprocess.run([CLI, "-i"]).returncode == 255


@pytest.mark.skipif(_have_gpu(), reason="a GPU is present")
def test_cli_fails_loudly_without_gpu(tmp_path):
    src = tmp_path / "in.u8"
    src.write_bytes(bytes(1024))
    r = subprocess.run([CLI, "-i", str(src), "-o", str(tmp_path / "out.f32")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_cli_reports_configuration_errors_before_it_needs_a_gpu(tmp_path):
    """the C host validates through fmd_create: argument errors come out as such (exit 2, message on
    stderr) on any machine; with valid arguments and no GPU it says that there is no CPU path"""
    import subprocess
    cli = os.path.join(ROOT, "demodulator_b200", "bin", "demodulator")
    if not os.path.exists(cli):
        pytest.skip("CLI not built")
    src = tmp_path / "in.u8"
    src.write_bytes(b"\x80\x7f" * 64)
    d = str(tmp_path).replace("-", "_")  # an -i/-o argument with a '-' means stdin/stdout (src/main.c:128)
    os.makedirs(d, exist_ok=True)
    inp, outp = os.path.join(d, "in.u8"), os.path.join(d, "out.f32")
    with open(inp, "wb") as f:
        f.write(src.read_bytes())

    def run(*flags):
        r = subprocess.run([cli, "-i", inp, "-o", outp, *flags], capture_output=True, text=True)
        return r.returncode, r.stderr
    assert "FMD_FLAG_S16" in run("-F", "s16")[1]
    assert "not an integer" in run("-Q", "-r100000", "-S48000")[1]
    assert "-q2" in run("-q2", "-b-6")[1]
    assert "-D without -L" in run("-D3")[1]
    rc, err = run("-F", "wav")
    assert rc == 1 and "f32 or s16" in err
    rc, err = run("-S96000", "-l12500")
    import torch
    if not torch.cuda.is_available():
        assert rc == 2 and "no CPU path" in err
