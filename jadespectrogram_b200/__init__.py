"""jadespectrogram_b200 -- B200-native (sm_100a) STFT -> dB -> colour-column engine behind JadeSpectrogram's API.

The product is the C-ABI shared library `libjade_gpu.so` (include/jade_gpu.h) plus the drop-in C++ classes in
include/Spectrogram.h / include/CColorpalette.h.  This package is the Python mirror of that interface used by the
tests and the benchmark.  There is no CPU fallback: without the built CUDA library the package raises.
"""
from ._capi import JadeConfig, load  # noqa: F401
from .engine import (Engine, JadeError, default_config, device_count, host_alloc, host_free,  # noqa: F401
                     render_batch_multi)
