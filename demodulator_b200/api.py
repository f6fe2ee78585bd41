"""ctypes mirror of include/fmdemod.h.

Field names follow the reference's ``consumerArgs`` (include/matrix.h:43-57): sampleRate,
lowpassIn, lowpassOut, epsilon, inFilterDegree, outFilterDegree, mode, bufSize.

The shared library is REQUIRED: importing works without it (so that CPU-only tooling can
import the package), but the first call that needs it raises ``FmdError`` loudly.  Nothing
here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CHUNK_BYTES = 262144
MAX_SECTIONS = 8
MAX_TAPS = 6
FLAG_EXACT = 1
FLAG_QUALITY = 2
FLAG_S16 = 4


class FmdError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("sample_rate", C.c_double), ("lowpass_in", C.c_double), ("lowpass_out", C.c_double),
                ("epsilon", C.c_double), ("in_degree", C.c_uint32), ("out_degree", C.c_uint32),
                ("mode", C.c_uint32), ("flags", C.c_uint32), ("chunk_bytes", C.c_uint64),
                ("input_rate", C.c_double)]


def lib_path() -> str:
    # FMD_LIB: kernel-tuning experiments point this at an alternative build of the same library
    return os.environ.get("FMD_LIB") or os.path.join(HERE, "lib", "libfmdemod.so")


_lib = None


def load_library() -> C.CDLL:
    """Load libfmdemod.so or fail loudly -- there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise FmdError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C demodulator_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u8p, f32p = C.c_void_p, C.c_void_p, C.c_void_p
    L.fmd_config_default.argtypes = [C.POINTER(_Config)]
    L.fmd_config_default.restype = None
    L.fmd_last_error.restype = C.c_char_p
    L.fmd_device_count.restype = C.c_int
    L.fmd_design_sos.argtypes = [C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, f32p]
    L.fmd_fold_taps.argtypes = [f32p, C.c_uint, C.c_int, f32p, f32p, C.POINTER(C.c_uint)]
    L.fmd_create.argtypes = [C.POINTER(_Config), C.c_int, C.POINTER(vp)]
    L.fmd_destroy.argtypes = [vp]
    L.fmd_destroy.restype = None
    L.fmd_get_config.argtypes = [vp, C.POINTER(_Config)]
    L.fmd_get_sos.argtypes = [vp, C.c_int, f32p]
    L.fmd_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.fmd_host_free.argtypes = [vp]
    L.fmd_host_register.argtypes = [vp, C.c_size_t]
    L.fmd_host_unregister.argtypes = [vp]
    L.fmd_out_count.argtypes = [vp, C.c_size_t]
    L.fmd_out_count.restype = C.c_size_t
    L.fmd_max_out_count.argtypes = [vp, C.c_size_t]
    L.fmd_max_out_count.restype = C.c_size_t
    L.fmd_reset.argtypes = [vp]
    L.fmd_demod_device.argtypes = [vp, u8p, C.c_size_t, f32p, vp]
    L.fmd_demod_host.argtypes = [vp, u8p, C.c_size_t, f32p, C.POINTER(C.c_size_t)]
    L.fmd_copy_host.argtypes = [vp, u8p, C.c_size_t, f32p]
    L.fmd_check.argtypes = [vp, vp]
    L.fmd_ring_create.argtypes = [vp, C.c_uint, C.c_size_t]
    L.fmd_submit.argtypes = [vp, C.c_uint, u8p, C.c_size_t, f32p]
    L.fmd_wait.argtypes = [vp, C.c_uint]
    L.fmd_warmup_for_sections.argtypes = [f32p, C.c_uint]
    L.fmd_warmup_for_sections.restype = C.c_uint64
    L.fmd_warmup_samples.argtypes = [vp]
    L.fmd_warmup_samples.restype = C.c_uint64
    L.fmd_segment_out_count.argtypes = [vp, C.c_uint64, C.c_size_t]
    L.fmd_segment_out_count.restype = C.c_size_t
    L.fmd_submit_segment.argtypes = [vp, C.c_uint, u8p, C.c_size_t, C.c_size_t, C.c_uint64, f32p]
    L.fmd_launch_count.argtypes = [vp]
    L.fmd_launch_count.restype = C.c_uint64
    L.fmd_profile_kernels.argtypes = [vp, C.c_int]
    L.fmd_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.fmd_time_device.argtypes = [vp, u8p, C.c_size_t, f32p, vp, C.c_int, C.POINTER(C.c_float)]
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc < 0:
        raise FmdError(f"libfmdemod error {rc}: {load_library().fmd_last_error().decode()}")


def device_count() -> int:
    n = load_library().fmd_device_count()
    return max(n, 0)


def design_sos(kind: int, degree: int, fc: float, fs: float, eps: float = 0.3) -> np.ndarray:
    """processFilterOption (src/matrix.c:25-80): rows {b0,b1,b2,a0,a1,a2} as float32."""
    rows = (degree >> 1) + (degree & 1)
    sos = np.zeros((max(rows, 1), 6), np.float32)
    _check(load_library().fmd_design_sos(kind, degree, fc, fs, eps, sos.ctypes.data))
    return sos[:rows]


def warmup_for_sections(sos: np.ndarray) -> int:
    """fmd_warmup_for_sections: samples after which a cascade started from a zero state has converged"""
    sos = np.ascontiguousarray(sos, np.float32)
    return int(load_library().fmd_warmup_for_sections(sos.ctypes.data, sos.shape[0]))


def fold_taps(sos: np.ndarray, bias: bool = True) -> tuple[np.ndarray, float]:
    sos = np.ascontiguousarray(sos, np.float32)
    taps = np.zeros(MAX_TAPS, np.float32)
    dc = np.zeros(1, np.float32)
    n = C.c_uint(0)
    _check(load_library().fmd_fold_taps(sos.ctypes.data, sos.shape[0], int(bias), taps.ctypes.data,
                                        dc.ctypes.data, C.byref(n)))
    return taps[:n.value].copy(), float(dc[0])


class Demodulator:
    """One context = one (GPU, channel) pair.  Arguments mirror consumerArgs / the CLI flags:

    sampleRate (-S), lowpassIn (-L), lowpassOut (-l), epsilon (-e, already divided by 10),
    inFilterDegree (-D), outFilterDegree (-d), mode (bit0 -m1 output Chebyshev, bit1 -m2 input
    Chebyshev, bits 2-3 the -q conditioner 0..3; the demod bit 0x10 is always on, as in
    src/main.c:112), bufSize (-b result).
    """

    def __init__(self, sampleRate: float = 125000.0, lowpassIn: float = 0.0, lowpassOut: float = 12500.0,
                 epsilon: float = 0.3, inFilterDegree: int = 0, outFilterDegree: int = 3, mode: int = 0,
                 bufSize: int = DEFAULT_CHUNK_BYTES, device: int = 0, exact: bool = False,
                 quality: bool = False, inputRate: float = 0.0, s16: bool = False) -> None:
        """quality=True selects the opt-in stream pipeline (FMD_FLAG_QUALITY: recursive filters with
        carried state, adjacent-sample discriminator, decimation by inputRate / sampleRate); s16=True
        (with quality) makes the outputs signed 16-bit samples (FMD_FLAG_S16: 1.0 = 32768, clipped)."""
        self._L = load_library()
        self._ctx = C.c_void_p()
        cfg = _Config(sampleRate, lowpassIn, lowpassOut, epsilon, inFilterDegree, outFilterDegree,
                      (mode & 0x3f) | 0x10, (FLAG_EXACT if exact else 0) | (FLAG_QUALITY if quality else 0) | (FLAG_S16 if s16 else 0),
                      bufSize, inputRate)
        _check(self._L.fmd_create(C.byref(cfg), device, C.byref(self._ctx)))
        eff = _Config()
        _check(self._L.fmd_get_config(self._ctx, C.byref(eff)))
        self.bufSize = int(eff.chunk_bytes)
        self.inFilterDegree = int(eff.in_degree)
        self.outFilterDegree = int(eff.out_degree)
        self.device = device
        self.out_dtype = np.int16 if s16 else np.float32

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.fmd_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "Demodulator":
        return self

    def __exit__(self, *a) -> None:
        self.close()

    # -- introspection ----------------------------------------------------------------------
    def sos(self, which: str = "out") -> np.ndarray:
        buf = np.zeros((MAX_SECTIONS, 6), np.float32)
        rows = self._L.fmd_get_sos(self._ctx, 1 if which == "in" else 0, buf.ctypes.data)
        _check(rows)
        return buf[:rows].copy()

    @property
    def launches(self) -> int:
        return int(self._L.fmd_launch_count(self._ctx))

    def out_count(self, nbytes: int) -> int:
        return int(self._L.fmd_out_count(self._ctx, nbytes))

    def max_out_count(self, nbytes: int) -> int:
        return int(self._L.fmd_max_out_count(self._ctx, nbytes))

    def reset(self) -> None:
        """quality mode: start a new stream (forget the carried filter state)."""
        _check(self._L.fmd_reset(self._ctx))

    # -- the hot path -----------------------------------------------------------------------
    def process(self, raw, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host floats out (fmd_demod_host: H2D -> kernels -> D2H)."""
        raw = np.ascontiguousarray(np.frombuffer(raw, np.uint8) if isinstance(raw, (bytes, bytearray)) else raw,
                                   np.uint8)
        n = self.out_count(raw.size)
        if out is None:
            out = np.empty(n, self.out_dtype)
        assert out.dtype == self.out_dtype and out.size >= n and out.flags.c_contiguous
        got = C.c_size_t(0)
        _check(self._L.fmd_demod_host(self._ctx, raw.ctypes.data, raw.size, out.ctypes.data, C.byref(got)))
        return out[:got.value]

    def process_ptr(self, in_ptr: int, nbytes: int, out_ptr: int) -> int:
        """fmd_demod_host on raw host addresses (pinned buffers owned by the caller)."""
        got = C.c_size_t(0)
        _check(self._L.fmd_demod_host(self._ctx, in_ptr, nbytes, out_ptr, C.byref(got)))
        return int(got.value)

    def copy_ptr(self, in_ptr: int, nbytes: int, out_ptr: int) -> None:
        """fmd_copy_host: the copies of process_ptr without the kernels (the e2e ceiling of the box)."""
        _check(self._L.fmd_copy_host(self._ctx, in_ptr, nbytes, out_ptr))

    def check(self, stream: int = 0) -> None:
        """fmd_check: synchronise `stream` and raise if the device flagged an error."""
        _check(self._L.fmd_check(self._ctx, stream))

    def process_device(self, d_in: int, nbytes: int, d_out: int, stream: int = 0) -> None:
        """Device pointers; asynchronous on `stream` (a cudaStream_t as int, 0 = the default stream)."""
        _check(self._L.fmd_demod_device(self._ctx, d_in, nbytes, d_out, stream))

    def time_device(self, d_in: int, nbytes: int, d_out: int, stream: int, reps: int) -> float:
        ms = C.c_float(0.0)
        _check(self._L.fmd_time_device(self._ctx, d_in, nbytes, d_out, stream, reps, C.byref(ms)))
        return float(ms.value)

    def profile_kernels(self, enable: bool) -> None:
        _check(self._L.fmd_profile_kernels(self._ctx, int(enable)))

    def kernel_time(self) -> tuple[float, int]:
        """(milliseconds, launches) of the fast kernel since the last call (CUDA events)."""
        ms, n = C.c_double(0.0), C.c_uint64(0)
        _check(self._L.fmd_kernel_time(self._ctx, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # -- streaming ring ---------------------------------------------------------------------
    def ring_create(self, slots: int, slot_bytes: int) -> None:
        _check(self._L.fmd_ring_create(self._ctx, slots, slot_bytes))

    def submit(self, slot: int, in_ptr: int, nbytes: int, out_ptr: int) -> None:
        _check(self._L.fmd_submit(self._ctx, slot, in_ptr, nbytes, out_ptr))

    def wait(self, slot: int) -> None:
        _check(self._L.fmd_wait(self._ctx, slot))

    # -- quality pipeline, one stream over several contexts / GPUs --------------------------
    def warmup_samples(self) -> int:
        return int(self._L.fmd_warmup_samples(self._ctx))

    def segment_out_count(self, first_sample: int, nsamples: int) -> int:
        return int(self._L.fmd_segment_out_count(self._ctx, first_sample, nsamples))

    def submit_segment(self, slot: int, in_ptr: int, warm_bytes: int, nbytes: int, first_sample: int, out_ptr: int) -> None:
        _check(self._L.fmd_submit_segment(self._ctx, slot, in_ptr, warm_bytes, nbytes, first_sample, out_ptr))


class PinnedBuffer:
    """fmd_host_alloc'd memory exposed as a numpy array (the allocateBuffer replacement)."""

    def __init__(self, nbytes: int) -> None:
        self._L = load_library()
        self._p = C.c_void_p()
        _check(self._L.fmd_host_alloc(C.byref(self._p), nbytes))
        self.nbytes = nbytes
        self.ptr = int(self._p.value)
        self.u8 = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def view(self, dtype) -> np.ndarray:
        return self.u8.view(dtype)

    def free(self) -> None:
        if self._p.value:
            self._L.fmd_host_free(self._p)
            self._p = C.c_void_p()
