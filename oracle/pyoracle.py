"""ctypes bindings for the CHECKERS under oracle/ -- test infrastructure only.

Two libraries, same call shapes:

* ``port()``  -> oracle/_build/libfm_oracle.so, this repo's plain-C restatement
  (fm_oracle.c).  Always buildable (gcc only).
* ``ref()``   -> oracle/_ref/libdemod_ref.so, the UNMODIFIED reference sources
  compiled in place (ref_shim.c).  Built in the authoring container, where
  /root/reference exists; travels to the GPU box as a prebuilt file.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
--impl reference) may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libfm_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libdemod_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "demodulator_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = ["port"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def sections(degree: int) -> int:
    return (degree >> 1) + (degree & 1)


@dataclass
class Config:
    """Mirror of the reference CLI's filter options (src/main.c:102-113)."""
    fs: float = 125000.0
    fc_in: float = 0.0
    fc_out: float = 12500.0
    eps: float = 0.3
    deg_in: int = 0
    deg_out: int = 3
    mode: int = 0  # bit0: output Chebyshev, bit1: input Chebyshev, bits 2-3: -q (0 or 3)


class _FmoConfig(C.Structure):
    _fields_ = [("fs", C.c_double), ("fc_in", C.c_double), ("fc_out", C.c_double),
                ("eps", C.c_double), ("deg_in", C.c_uint), ("deg_out", C.c_uint),
                ("mode", C.c_uint)]


class Port:
    """oracle/fm_oracle.c"""
    kind = "port"

    def __init__(self) -> None:
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = L = C.CDLL(PORT_SO)
        L.fmo_design_sos.argtypes = [C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, _f32p]
        L.fmo_design_sos.restype = None
        L.fmo_shift.argtypes = [_u8p, C.c_size_t, _f32p]
        L.fmo_shift.restype = None
        L.fmo_discriminate.argtypes = [_f32p, C.c_size_t, _f32p]
        L.fmo_discriminate.restype = None
        L.fmo_normalize.argtypes = [_u8p, C.c_size_t, _f32p]
        L.fmo_normalize.restype = None
        L.fmo_sweep_lane.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _f32p, C.c_int]
        L.fmo_sweep_lane.restype = None
        L.fmo_process_chunks.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.POINTER(_FmoConfig), _f32p]
        L.fmo_process_chunks.restype = C.c_size_t
        L.fmo_quality_state_size.restype = C.c_size_t
        L.fmo_quality_init.argtypes = [C.c_void_p, C.POINTER(_FmoConfig), C.c_double]
        L.fmo_quality_init.restype = C.c_uint
        L.fmo_quality_run.argtypes = [C.c_void_p, _u8p, C.c_size_t, _f32p]
        L.fmo_quality_run.restype = C.c_size_t
        _f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.fmo_quality_stages.argtypes = [C.c_void_p, _u8p, C.c_size_t, _f64p, _f64p, _f64p]
        L.fmo_quality_stages.restype = None

    def quality(self, cfg: "Config", input_rate: float) -> "QualityStream":
        """Streaming state of the opt-in quality pipeline (fm_quality_oracle.c)."""
        return QualityStream(self.lib, cfg, input_rate)

    def design(self, kind: int, degree: int, fc: float, fs: float, eps: float = 0.3) -> np.ndarray:
        sos = np.zeros((sections(degree), 6), np.float32)
        self.lib.fmo_design_sos(kind, degree, fc, fs, eps, sos)
        return sos

    def shift(self, raw: np.ndarray) -> np.ndarray:
        out = np.empty(raw.size, np.float32)
        self.lib.fmo_shift(np.ascontiguousarray(raw, np.uint8), raw.size, out)
        return out

    def normalize(self, raw: np.ndarray) -> np.ndarray:
        out = np.empty(raw.size, np.float32)
        self.lib.fmo_normalize(np.ascontiguousarray(raw, np.uint8), raw.size, out)
        return out

    def discriminate(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, np.float32)
        out = np.empty(iq.size // 4, np.float32)
        self.lib.fmo_discriminate(iq, iq.size, out)
        return out

    def sweep_real(self, x: np.ndarray, sos: np.ndarray, init: np.ndarray | None = None) -> np.ndarray:
        """applyFilter on n=len(x) samples; returns the n + 2L floats of the y region."""
        L = sos.shape[0]
        n = x.size
        xe = np.zeros(n + 2 * L, np.float32)
        xe[:n] = x
        y = np.zeros(n + 2 * L, np.float32)
        if init is not None:
            y[:init.size] = init
        self.lib.fmo_sweep_lane(xe, y, n, L, np.ascontiguousarray(sos, np.float32), 1)
        return y

    def sweep_complex(self, x: np.ndarray, sos: np.ndarray) -> np.ndarray:
        """applyComplexFilter on interleaved IQ floats with ZERO extension; returns n floats."""
        L = sos.shape[0]
        nc = x.size // 2
        out = np.empty(2 * nc, np.float32)
        for q in (0, 1):
            xe = np.zeros(nc + 2 * L, np.float32)
            xe[:nc] = x[q::2]
            y = np.zeros(nc + 2 * L, np.float32)
            self.lib.fmo_sweep_lane(xe, y, nc, L, np.ascontiguousarray(sos, np.float32), int(q == 0))
            out[q::2] = y[:nc]
        return out

    def run(self, raw: np.ndarray, cfg: Config, chunk_bytes: int = 262144) -> np.ndarray:
        raw = np.ascontiguousarray(raw, np.uint8)
        nchunks = raw.size // chunk_bytes
        out = np.empty(nchunks * (chunk_bytes // 4), np.float32)
        c = _FmoConfig(cfg.fs, cfg.fc_in, cfg.fc_out, cfg.eps, cfg.deg_in, cfg.deg_out, cfg.mode)
        n = self.lib.fmo_process_chunks(raw, nchunks, chunk_bytes, C.byref(c), out)
        if n != out.size:
            raise ValueError("oracle port rejected the configuration")
        return out


class QualityStream:
    """fm_quality_oracle.c: causal IIR cascades with carried state, adjacent-sample discriminator,
    decimation by M = input_rate / cfg.fs.  Double precision, serial."""

    def __init__(self, lib, cfg: Config, input_rate: float) -> None:
        self.lib = lib
        self._st = C.create_string_buffer(lib.fmo_quality_state_size())
        c = _FmoConfig(cfg.fs, cfg.fc_in, cfg.fc_out, cfg.eps, cfg.deg_in, cfg.deg_out, cfg.mode)
        self.decim = int(lib.fmo_quality_init(self._st, C.byref(c), input_rate))
        if self.decim == 0:
            raise ValueError("quality oracle rejected the configuration")

    def run(self, raw: np.ndarray) -> np.ndarray:
        raw = np.ascontiguousarray(raw, np.uint8)
        n = raw.size // 2
        out = np.empty(n // self.decim + 2, np.float32)
        got = self.lib.fmo_quality_run(self._st, raw, n, out)
        return out[:got].copy()

    def stages(self, raw: np.ndarray):
        raw = np.ascontiguousarray(raw, np.uint8)
        n = raw.size // 2
        z, x, v = np.empty(2 * n), np.empty(n), np.empty(n)
        self.lib.fmo_quality_stages(self._st, raw, n, z, x, v)
        return z, x, v


class Ref:
    """oracle/_ref/libdemod_ref.so -- the reference's own code."""
    kind = "reference"

    def __init__(self) -> None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(
                f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = L = C.CDLL(REF_SO)
        L.ref_design.argtypes = [C.c_uint, C.c_size_t, C.c_double, C.c_double, C.c_double, _f32p]
        L.ref_design.restype = None
        L.ref_shift_origin.argtypes = [_u8p, C.c_size_t, _f32p]
        L.ref_shift_origin.restype = None
        L.ref_fm_demod.argtypes = [_f32p, C.c_size_t, _f32p]
        L.ref_fm_demod.restype = None
        L.ref_normalize_input.argtypes = [_u8p, C.c_size_t, _f32p]
        L.ref_normalize_input.restype = None
        L.ref_apply_filter.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _f32p]
        L.ref_apply_filter.restype = None
        L.ref_apply_complex_filter.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _f32p]
        L.ref_apply_complex_filter.restype = None
        L.ref_run_chunks.argtypes = [_u8p, C.c_size_t, C.c_size_t, C.c_uint, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_size_t, C.c_size_t, _f32p]
        L.ref_run_chunks.restype = C.c_size_t

    def design(self, kind: int, degree: int, fc: float, fs: float, eps: float = 0.3) -> np.ndarray:
        sos = np.zeros((sections(degree), 6), np.float32)
        self.lib.ref_design(kind, degree, fc, fs, eps, sos)
        return sos

    def shift(self, raw: np.ndarray) -> np.ndarray:
        out = np.empty(raw.size, np.float32)
        self.lib.ref_shift_origin(np.ascontiguousarray(raw, np.uint8), raw.size, out)
        return out

    def normalize(self, raw: np.ndarray) -> np.ndarray:
        out = np.empty(raw.size, np.float32)
        self.lib.ref_normalize_input(np.ascontiguousarray(raw, np.uint8), raw.size, out)
        return out

    def discriminate(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, np.float32)
        out = np.empty(iq.size // 4, np.float32)
        self.lib.ref_fm_demod(iq, iq.size, out)
        return out

    def sweep_real(self, x: np.ndarray, sos: np.ndarray, init: np.ndarray | None = None) -> np.ndarray:
        L = sos.shape[0]
        n = x.size
        xe = np.zeros(n + 2 * L, np.float32)
        xe[:n] = x
        y = np.zeros(n + 2 * L, np.float32)
        if init is not None:
            y[:init.size] = init
        self.lib.ref_apply_filter(xe, y, n, L, np.ascontiguousarray(sos, np.float32))
        return y

    def sweep_complex(self, x: np.ndarray, sos: np.ndarray) -> np.ndarray:
        L = sos.shape[0]
        n = x.size
        xe = np.zeros(n + 4 * L, np.float32)
        xe[:n] = x
        y = np.zeros(n + 4 * L, np.float32)
        self.lib.ref_apply_complex_filter(xe, y, n, L, np.ascontiguousarray(sos, np.float32))
        return y[:n].copy()

    def run(self, raw: np.ndarray, cfg: Config, chunk_bytes: int = 262144) -> np.ndarray:
        if ((cfg.mode >> 2) & 3) == 1:
            # -q1 keeps its state in a function-local static of the reference: the shim cannot start it
            # afresh per call.  The port answers; it is pinned bit for bit against the reference CLI's
            # own -q1 outputs (tests/golden/q1*.npz).
            return port().run(raw, cfg, chunk_bytes)
        raw = np.ascontiguousarray(raw, np.uint8)
        nchunks = raw.size // chunk_bytes
        out = np.empty(nchunks * (chunk_bytes // 4), np.float32)
        n = self.lib.ref_run_chunks(raw, nchunks, chunk_bytes, cfg.mode, cfg.fs, cfg.fc_in,
                                    cfg.fc_out, cfg.eps, cfg.deg_in, cfg.deg_out, out)
        if n != out.size:
            raise ValueError("reference shim rejected the configuration")
        return out


_port: Port | None = None
_ref: Ref | None = None


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref() -> Ref:
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def best():
    """The real reference when its prebuilt library is present, else the port."""
    return ref() if have_ref() else port()
