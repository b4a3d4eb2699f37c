"""Python mirror of the C ABI (include/jade_gpu.h): a thin `Engine` handle used by tests and bench.py.

All compute goes through libjade_gpu.so (CUDA, sm_100a).  Nothing here computes a spectrogram on the CPU; the module
raises `JadeError` if the library is missing, no GPU is present, or a call fails.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import EMIT, MIX, PAL, PIX, ROWS, SYNTH, WIN, JadeConfig


class JadeError(RuntimeError):
    pass


def _enum(table, v):
    return table[v] if isinstance(v, str) else int(v)


def default_config(**kw):
    """jade_config_default (the plugin's live defaults) with keyword overrides; `feed_percent=` applies the
    reference's 100/50/25/10 % rule after fft_size is set."""
    lib = _capi.load()
    c = JadeConfig()
    lib.jade_config_default(C.byref(c))
    feed = kw.pop("feed_percent", None)
    enums = dict(window=WIN, mix_mode=MIX, row_map=ROWS, pixel_format=PIX, emit_mode=EMIT)
    explicit_hop = "hop" in kw
    for k, v in kw.items():
        if k in enums:
            v = _enum(enums[k], v)
        if not hasattr(c, k):
            raise TypeError(f"jade_config has no field {k!r}")
        setattr(c, k, v)
    if feed is not None:
        if lib.jade_config_set_feed_percent(C.byref(c), int(feed)) != 0:
            raise JadeError(f"bad feed percent {feed}")
    elif explicit_hop:
        if "frames_per_block" not in kw:
            c.frames_per_block = 1
        if "block_stride" not in kw:
            c.block_stride = c.hop * c.frames_per_block
        if "emit_mode" not in kw:
            c.emit_mode = EMIT["hop"]
    else:
        lib.jade_config_set_feed_percent(C.byref(c), 50)
    return c


def display_freq_clamp(fs, min_hz, max_hz):
    """Slider clamp rules of SpectrogramComponent::paint (Spectrogram.cpp:441-453); host-only."""
    a, b = C.c_float(min_hz), C.c_float(max_hz)
    if _capi.load().jade_display_freq_clamp(float(fs), C.byref(a), C.byref(b)) != 0:
        raise JadeError("bad arguments")
    return a.value, b.value


def axis_ticks(kind, lo, hi, comp_height=550, scale=1.0, menu_height=20, text_height=20, nticks=11):
    """Tick values / label-box positions of the frequency ("freq") or colourbar ("color") axis (Spectrogram.cpp:466-545);
    host-only.  Returns (values float32[nticks], y int32[nticks], labels)."""
    lib = _capi.load()
    t = (_capi.AxisTick * nticks)()
    fn = lib.jade_freq_axis_ticks if kind == "freq" else lib.jade_color_axis_ticks
    if fn(float(lo), float(hi), int(comp_height), float(scale), int(menu_height), int(text_height), int(nticks), t) != 0:
        raise JadeError("bad arguments")
    return (np.array([x.value for x in t], np.float32), np.array([x.y for x in t], np.int32), [x.label.decode() for x in t])


def colorbar_height(comp_height=550, scale=1.0, menu_height=20):
    return _capi.load().jade_colorbar_height(int(comp_height), float(scale), int(menu_height))


class Engine:
    """One jade_engine (one GPU)."""

    def __init__(self, device=0, config=None, **cfg_kw):
        self.lib = _capi.load()
        h = C.c_void_p()
        rc = self.lib.jade_create(int(device), C.byref(h))
        if rc != 0:
            raise JadeError(f"jade_create failed ({rc}): {self.lib.jade_last_error(None).decode()}")
        self.h = h
        self.device = device
        if config is not None or cfg_kw:
            self.configure(config, **cfg_kw)

    # -- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise JadeError(f"libjade_gpu error {rc}: {self.lib.jade_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.jade_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def configure(self, config=None, **kw):
        c = config if config is not None else default_config(**kw)
        self._ck(self.lib.jade_configure(self.h, C.byref(c)))
        self.cfg = self.get_config()
        self.N = self.cfg.fft_size
        self.B = self.N // 2 + 1
        self.R = self.cfg.rows
        self.W = self.cfg.ring_columns
        return self

    def get_config(self):
        c = JadeConfig()
        self._ck(self.lib.jade_get_config(self.h, C.byref(c)))
        return c

    def set_pause(self, on):
        self._ck(self.lib.jade_set_pause(self.h, int(bool(on))))

    def set_window(self, w):
        self._ck(self.lib.jade_set_window(self.h, _enum(WIN, w)))

    def get_window(self):
        out = np.empty(self.N, np.float32)
        self._ck(self.lib.jade_get_window(self.h, out.ctypes.data, self.N))
        return out

    def reset(self):
        self._ck(self.lib.jade_reset(self.h))

    # -- palette
    def set_palette(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.int32)
        self._ck(self.lib.jade_set_palette(self.h, rgb.ctypes.data, rgb.size))

    def set_palette_scheme(self, scheme, n=256, invert=False):
        self._ck(self.lib.jade_set_palette_scheme(self.h, _enum(PAL, scheme), int(n), int(bool(invert))))

    def set_value_range(self, mn, mx):
        self._ck(self.lib.jade_set_value_range(self.h, float(mn), float(mx)))

    def get_value_range(self):
        a, b, m = C.c_float(), C.c_float(), C.c_float()
        self._ck(self.lib.jade_get_value_range(self.h, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, m.value

    def colorbar(self, height, ramp_min=-50.0, ramp_max=50.0):
        """The colourbar column of SpectrogramComponent::paint (Spectrogram.cpp:511-521), top pixel first."""
        out = np.empty(int(height), np.uint32)
        self._ck(self.lib.jade_colorbar(self.h, int(height), float(ramp_min), float(ramp_max), out.ctypes.data))
        return out

    def lookup_color(self, v):
        out = C.c_int32()
        self._ck(self.lib.jade_lookup_color(self.h, float(v), C.byref(out)))
        return out.value

    # -- streaming
    def push(self, planar):
        """planar [channels][n] float32 host array."""
        planar = np.ascontiguousarray(planar, np.float32)
        if planar.ndim == 1:
            planar = planar[None, :]
        ch, n = planar.shape
        ptrs = (C.c_void_p * ch)(*[planar[i].ctypes.data for i in range(ch)])
        self._ck(self.lib.jade_push_samples(self.h, ptrs, ch, n))

    def fetch(self, max_cols=None, want_db=True):
        if max_cols is None:
            max_cols = self.W
        pix = np.empty((max_cols, self.R), np.uint32)
        db = np.empty((max_cols, self.B), np.float32) if want_db else None
        n = C.c_int(0)
        first = C.c_int64(0)
        self._ck(self.lib.jade_fetch_columns(self.h, pix.ctypes.data, db.ctypes.data if want_db else None, max_cols,
                                             C.byref(n), C.byref(first)))
        return pix[:n.value], (db[:n.value] if want_db else None), first.value

    def ring_info(self):
        w, r, b, t = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._ck(self.lib.jade_ring_info(self.h, C.byref(w), C.byref(r), C.byref(b), C.byref(t)))
        return w.value, r.value, b.value, t.value

    def recolor_ring(self):
        pix = np.empty((self.W, self.R), np.uint32)
        self._ck(self.lib.jade_recolor_ring(self.h, pix.ctypes.data))
        return pix

    def read_ring_db(self):
        db = np.empty((self.W, self.B), np.float32)
        self._ck(self.lib.jade_read_ring_db(self.h, db.ctypes.data))
        return db

    # -- batch
    def columns_for(self, nsamples):
        return int(self.lib.jade_columns_for(self.h, int(nsamples)))

    def render_batch(self, samples, first_col=0, ncols=None, want_pix=True, want_db=False, out_pix=None, out_db=None):
        """samples [nstreams][channels][nsamples] (host) -> pixels [nstreams][ncols][rows], db [nstreams][ncols][bins]."""
        samples = np.asarray(samples)
        if samples.dtype != np.float32 or not samples.flags["C_CONTIGUOUS"]:
            samples = np.ascontiguousarray(samples, np.float32)
        if samples.ndim == 2:
            samples = samples[None]
        ns, ch, n = samples.shape
        if ch != self.cfg.channels:
            raise JadeError(f"samples have {ch} channels, engine configured for {self.cfg.channels}")
        if ncols is None:
            ncols = self.columns_for(n) - first_col
        pix = out_pix if out_pix is not None else (np.empty((ns, ncols, self.R), np.uint32) if want_pix else None)
        db = out_db if out_db is not None else (np.empty((ns, ncols, self.B), np.float32) if want_db else None)
        self._ck(self.lib.jade_render_batch(self.h, samples.ctypes.data, ns, n, int(first_col), int(ncols),
                                            pix.ctypes.data if pix is not None else None,
                                            db.ctypes.data if db is not None else None))
        return pix, db

    def render_device(self, d_samples_ptr, nstreams, nsamples, stream_stride, channel_stride, first_col, ncols,
                      d_pix_ptr, d_db_ptr=None, cuda_stream=None):
        self._ck(self.lib.jade_render_device(self.h, d_samples_ptr, int(nstreams), int(nsamples), int(stream_stride),
                                             int(channel_stride), int(first_col), int(ncols), d_pix_ptr, d_db_ptr,
                                             cuda_stream))

    def synth_device(self, d_out_ptr, nstreams, channels, nsamples, stream_stride, channel_stride, kind="mix", seed=20240601,
                     cuda_stream=None):
        self._ck(self.lib.jade_synth_device(self.h, d_out_ptr, int(nstreams), int(channels), int(nsamples),
                                            int(stream_stride), int(channel_stride), _enum(SYNTH, kind), int(seed),
                                            cuda_stream))

    def sync(self):
        self._ck(self.lib.jade_sync(self.h))

    @property
    def kernel_launches(self):
        return int(self.lib.jade_kernel_launches(self.h))

    @property
    def kernel_name(self):
        return self.lib.jade_kernel_name(self.h).decode()

    def last_kernel_seconds(self):
        return float(self.lib.jade_last_kernel_seconds(self.h))


_PINNED = {}


def device_count():
    """CUDA devices visible to libjade_gpu.so (0 without a GPU)."""
    return int(_capi.load().jade_device_count())


def render_batch_multi(engines, samples, first_col=0, ncols=None, want_pix=True, want_db=False):
    """jade_render_batch_multi: one identically configured engine per GPU; streams are range-partitioned (a single stream:
    its columns, each GPU re-reading the N - hop input halo), no inter-GPU exchange.  Returns (pixels, db) like render_batch."""
    samples = np.ascontiguousarray(samples, np.float32)
    if samples.ndim == 2:
        samples = samples[None]
    ns, ch, n = samples.shape
    e0 = engines[0]
    if ncols is None:
        ncols = e0.columns_for(n) - first_col
    pix = np.empty((ns, ncols, e0.R), np.uint32) if want_pix else None
    db = np.empty((ns, ncols, e0.B), np.float32) if want_db else None
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rc = e0.lib.jade_render_batch_multi(arr, len(engines), samples.ctypes.data, ns, n, int(first_col), int(ncols),
                                        pix.ctypes.data if pix is not None else None, db.ctypes.data if db is not None else None)
    if rc != 0:
        raise JadeError(f"jade_render_batch_multi failed ({rc}): {e0.lib.jade_last_error(None).decode()}")
    return pix, db


def host_alloc(shape, dtype):
    """numpy array backed by page-locked memory from jade_host_alloc; release it with host_free(arr)."""
    lib = _capi.load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    p = lib.jade_host_alloc(nbytes)
    if not p:
        raise JadeError("jade_host_alloc failed: " + lib.jade_last_error(None).decode())
    buf = (C.c_char * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


def host_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        _capi.load().jade_host_free(p)
