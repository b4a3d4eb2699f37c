"""The C-ABI library loads without a GPU, exports every symbol include/jade_gpu.h declares, and refuses to compute
without a device (no CPU fallback).  The drop-in headers compile against it."""
import ctypes as C
import pathlib
import re
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _declared():
    hdr = (ROOT / "include" / "jade_gpu.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(jade_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound():
    from jadespectrogram_b200 import _capi
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jade_gpu.h but not exported by libjade_gpu.so"
    assert sorted(_capi.SYMBOLS) == names, "jadespectrogram_b200/_capi.py and include/jade_gpu.h disagree"
    assert lib.jade_abi_version() == 1


def test_exports_are_c_linkage_only():
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "jadespectrogram_b200" / "libjade_gpu.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    jade = [s for s in exported if s.startswith("jade_")]
    assert set(_declared()) <= set(jade)


def test_config_defaults_are_the_plugins():
    from jadespectrogram_b200 import _capi
    lib = _capi.load()
    c = _capi.JadeConfig()
    assert lib.jade_config_default(C.byref(c)) == 0
    # PluginProcessor.cpp:13,102-114: 2048-point FFT, 50 % feed, 10 s; Spectrogram.cpp:17,21-22: 2 ch, AbsMean, Hann
    assert (c.fft_size, c.hop, c.frames_per_block, c.block_stride, c.channels, c.window, c.mix_mode) == (2048, 1024, 2, 2048, 2, 1, 0)
    assert c.memory_time_s == 10.0 and c.sample_rate == 48000.0 and c.flip_y == 1
    for pct, hop, fb in ((100, 2048, 1), (50, 1024, 2), (25, 512, 4), (10, 205, 10)):
        assert lib.jade_config_set_feed_percent(C.byref(c), pct) == 0
        assert (c.hop, c.frames_per_block) == (hop, fb)
    assert lib.jade_config_set_feed_percent(C.byref(c), 33) != 0


def test_no_cpu_fallback_without_a_device():
    import torch
    from jadespectrogram_b200 import Engine, JadeError, _capi
    lib = _capi.load()
    if torch.cuda.is_available():
        assert lib.jade_device_count() >= 1
        return
    assert lib.jade_device_count() == 0
    h = C.c_void_p()
    assert lib.jade_create(0, C.byref(h)) == _capi.JADE_ERR_NOGPU and not h.value
    assert b"no CPU fallback" in lib.jade_last_error(None)
    try:
        Engine(0)
    except JadeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("Engine() must fail without a GPU")


def test_dropin_class_fails_loudly_without_a_device():
    import torch
    so = ROOT / "jadespectrogram_b200" / "libjade_dropin_shim.so"
    L = C.CDLL(str(so))
    L.jd_spec_create.restype = C.c_void_p
    L.jd_spec_error.restype = C.c_char_p
    for f in ("jd_spec_destroy", "jd_spec_ok", "jd_spec_error", "jd_spec_spectrum_size", "jd_spec_memory_size"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.jd_spec_process_block.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    s = L.jd_spec_create()
    # geometry getters mirror the reference constructor defaults either way
    assert (L.jd_spec_spectrum_size(s), L.jd_spec_memory_size(s)) == (513, 47)
    if not torch.cuda.is_available():
        assert L.jd_spec_ok(s) == 0 and b"no CPU fallback" in L.jd_spec_error(s)
        x = np.zeros((2, 1024), np.float32)
        assert L.jd_spec_process_block(s, x.ctypes.data, 2, 1024) == -1
    L.jd_spec_destroy(s)
