"""ctypes binding of the C ABI in include/jade_gpu.h (libjade_gpu.so, built in-tree by jadespectrogram_b200/csrc).

The library is the product; this module only declares its symbols.  It fails loudly when the shared library is
missing or has no usable GPU -- there is no CPU path.
"""
import ctypes as C
import os
import pathlib

PKG_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("JADE_GPU_LIB", PKG_DIR / "libjade_gpu.so"))  # override: kernel-variant experiments

# enums (include/jade_gpu.h)
MIX = dict(absmean=0, max=1, min=2, left=3, right=4)
WIN = dict(rect=0, hann=1, hamming=2, blackmanharris=3, flattop=4, hannpoisson=5)
PAL = dict(mono=0, bw=1, hot=2, rainbow=3, viridis=4, plasma=5, jade=6)
ROWS = dict(identity=0, linear_crop=1, log_maxpool=2)
PIX = dict(argb32=0, rgba8=1)
EMIT = dict(hop=0, block=1)
SYNTH = dict(sweep=0, noise=1, mix=2)

JADE_OK, JADE_ERR_ARG, JADE_ERR_CUDA, JADE_ERR_STATE, JADE_ERR_NOGPU = 0, -1, -2, -3, -4


class JadeConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_float), ("fft_size", C.c_int32), ("hop", C.c_int32), ("frames_per_block", C.c_int32),
        ("block_stride", C.c_int32), ("preroll", C.c_int32), ("emit_mode", C.c_int32), ("window", C.c_int32),
        ("channels", C.c_int32), ("mix_mode", C.c_int32), ("row_map", C.c_int32), ("rows", C.c_int32),
        ("fmin", C.c_float), ("fmax", C.c_float), ("flip_y", C.c_int32), ("pixel_format", C.c_int32),
        ("power_scale", C.c_float), ("memory_time_s", C.c_float), ("ring_columns", C.c_int32),
        ("db_precise", C.c_int32), ("max_push", C.c_int32),
    ]


class AxisTick(C.Structure):
    _fields_ = [("value", C.c_float), ("y", C.c_int32), ("label", C.c_char * 16)]


# every symbol include/jade_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_CFG = C.POINTER(JadeConfig)
_I = C.c_int
_I64 = C.c_int64
_F = C.c_float
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)
SYMBOLS = {
    "jade_abi_version": (_I, []),
    "jade_device_count": (_I, []),
    "jade_create": (_I, [_I, _PP]),
    "jade_destroy": (_I, [_P]),
    "jade_last_error": (C.c_char_p, [_P]),
    "jade_config_default": (_I, [_CFG]),
    "jade_config_set_feed_percent": (_I, [_CFG, _I]),
    "jade_configure": (_I, [_P, _CFG]),
    "jade_get_config": (_I, [_P, _CFG]),
    "jade_set_pause": (_I, [_P, _I]),
    "jade_set_window": (_I, [_P, _I]),
    "jade_get_window": (_I, [_P, _P, _I]),
    "jade_window_build": (_I, [_I, _I, _P]),
    "jade_reset": (_I, [_P]),
    "jade_palette_build": (_I, [_I, _I, _I, _P]),
    "jade_set_palette": (_I, [_P, _P, _I]),
    "jade_set_palette_scheme": (_I, [_P, _I, _I, _I]),
    "jade_set_value_range": (_I, [_P, _F, _F]),
    "jade_get_value_range": (_I, [_P, _FP, _FP, _FP]),
    "jade_lookup_color": (_I, [_P, _F, C.POINTER(C.c_int32)]),
    "jade_linear_crop": (_I, [_F, _I, _F, _F, _IP, _IP]),
    "jade_display_freq_clamp": (_I, [_F, _FP, _FP]),
    "jade_freq_axis_ticks": (_I, [_F, _F, _I, _F, _I, _I, _I, _P]),
    "jade_color_axis_ticks": (_I, [_F, _F, _I, _F, _I, _I, _I, _P]),
    "jade_colorbar_height": (_I, [_I, _F, _I]),
    "jade_colorbar": (_I, [_P, _I, _F, _F, _P]),
    "jade_log_rows": (_I, [_F, _I, _I, _F, _F, _P, _P]),
    "jade_push_samples": (_I, [_P, C.POINTER(_P), _I, _I]),
    "jade_fetch_columns": (_I, [_P, _P, _P, _I, _IP, C.POINTER(_I64)]),
    "jade_ring_info": (_I, [_P, _IP, _IP, _IP, C.POINTER(_I64)]),
    "jade_recolor_ring": (_I, [_P, _P]),
    "jade_read_ring_db": (_I, [_P, _P]),
    "jade_view_create": (_I, [_P, _PP]),
    "jade_view_destroy": (_I, [_P]),
    "jade_view_set_running": (_I, [_P, _I]),
    "jade_view_set_value_range": (_I, [_P, _F, _F]),
    "jade_view_invalidate": (_I, [_P]),
    "jade_view_tick": (_I, [_P, _IP]),
    "jade_view_image": (_I, [_P, C.POINTER(C.POINTER(C.c_uint32)), _IP, _IP]),
    "jade_columns_for": (_I64, [_P, _I64]),
    "jade_render_batch": (_I, [_P, _P, _I, _I64, _I64, _I64, _P, _P]),
    "jade_render_batch_multi": (_I, [C.POINTER(_P), _I, _P, _I, _I64, _I64, _I64, _P, _P]),
    "jade_render_device": (_I, [_P, _P, _I, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "jade_sync": (_I, [_P]),
    "jade_synth_device": (_I, [_P, _P, _I, _I, _I64, _I64, _I64, _I, C.c_uint64, _P]),
    "jade_kernel_launches": (_I64, [_P]),
    "jade_kernel_name": (C.c_char_p, [_P]),
    "jade_last_kernel_seconds": (C.c_double, [_P]),
    "jade_host_alloc": (_P, [C.c_size_t]),
    "jade_host_free": (_I, [_P]),
}

_lib = None


def load():
    """Load libjade_gpu.so and declare its prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C jadespectrogram_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            f = getattr(lib, name)  # AttributeError if the symbol is not exported
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib
