"""demodulator_b200 -- B200-native FM demodulation hot path (drop-in for peads/demodulator's
processMatrix path).  The product is the C-ABI library demodulator_b200/lib/libfmdemod.so
(include/fmdemod.h) and the C CLI demodulator_b200/bin/demodulator; this package is the thin
Python mirror used by the tests and the bench.  There is no CPU fallback."""
from .api import (Demodulator, FmdError, design_sos, device_count, fold_taps, lib_path, load_library,  # noqa: F401
                  DEFAULT_CHUNK_BYTES)
