"""icer_compression_b200 -- Blackwell-native ICER encoder/decoder behind lib_icer's C API.

The product is the C-ABI shared library `lib/libicer_b200.so` (hand-written sm_100a CUDA kernels + C++ host
orchestration; sources in `csrc/`, headers in `/include`).  This Python package is only the host-side mirror of
that ABI used by the tests and bench.py: it loads the library with ctypes and exposes the reference's entry
points under their own names plus the batch extension.  There is no Python or CPU codec path here: if the
library is missing or no CUDA device is usable every call raises.
"""
from .api import (  # noqa: F401
    IcerB200, IcerMulti, LegacyAPI, IcerError, device_count, lib_path, load_library, MEM_HOST, MEM_DEVICE, PIX_U16, PIX_U8, PIX_RGB888, PIX_NATIVE_U8, u8_interleave_map,
    ICER_RESULT_OK, ICER_INTEGER_OVERFLOW, ICER_OUTPUT_BUF_TOO_SMALL, ICER_TOO_MANY_SEGMENTS, ICER_TOO_MANY_STAGES,
    ICER_BYTE_QUOTA_EXCEEDED, ICER_FATAL_ERROR, FILTERS,
)
