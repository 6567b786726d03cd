"""fusion_power_video_b200 -- B200 (sm_100a) pre-entropy transform for
fusion-power-video streams.

The product is the CUDA shared library ``lib/libfpv_b200.so`` (C ABI declared
in ``include/fpv_b200.h``) and the C++ host layer built on it
(``csrc/host``).  This Python package is a thin ctypes binding over that C ABI
used by the tests and ``bench.py``; it contains no compute of its own and no
CPU fallback -- importing :mod:`fusion_power_video_b200.binding` raises if the
library has not been built.
"""
from .binding import (  # noqa: F401
    Context,
    FpvError,
    ENC_DEFAULT,
    ENC_GENERIC,
    ENC_NO_DELTA,
    DEC_DEFAULT,
    DEC_UNEXTRACT,
    FLAG_USE_DELTA,
    FLAG_USE_CG,
    FLAG_NO_LOW_BYTES,
    device_count,
    lib,
    lib_path,
    version,
)
