"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and of the compiled reference (oracle/_ref/libjade_ref.so).

TEST INFRASTRUCTURE: imported only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
import ctypes as C
import os
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

MIX = dict(absmean=0, max=1, min=2, left=3, right=4)
WIN = dict(rect=0, hann=1, hamming=2, blackmanharris=3, flattop=4, hannpoisson=5)
FEED = dict(p100=0, p50=1, p25=2, p10=3)
PAL = dict(mono=0, bw=1, hot=2, rainbow=3, viridis=4, plasma=5, jade=6)

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


class BatchCfg(C.Structure):
    _fields_ = [("fs", C.c_float), ("fft_size", C.c_int), ("hop", C.c_int), ("window", C.c_int),
                ("channels", C.c_int), ("mix_mode", C.c_int), ("palette_scheme", C.c_int),
                ("palette_size", C.c_int), ("palette_invert", C.c_int), ("min_db", C.c_float),
                ("max_db", C.c_float), ("use_double_fft", C.c_int)]


def build_oracle():
    """(Re)build liboracle.so (and _ref when /root/reference is present). Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR)], check=True, stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            build_oracle()
        L = C.CDLL(str(so))
        L.jo_window.argtypes = [C.c_int, C.c_int, f32p]
        L.jo_power_f32.argtypes = [f32p, C.c_int, f32p]
        L.jo_power_f64.argtypes = [f32p, C.c_int, f64p]
        L.jo_db.argtypes = [C.c_float]
        L.jo_db.restype = C.c_float
        L.jo_spec_create.restype = C.c_void_p
        for n, a in [("destroy", []), ("set_samplerate", [C.c_float]), ("set_channels", [C.c_size_t]),
                     ("set_fftsize", [C.c_size_t]), ("set_closest_fftsize_ms", [C.c_float]),
                     ("set_memory_time_s", [C.c_float]), ("set_feed_percent", [C.c_int]), ("set_pause", [C.c_int]),
                     ("set_window", [C.c_int]), ("set_mix_mode", [C.c_int]), ("set_fft_double", [C.c_int])]:
            f = getattr(L, "jo_spec_" + n)
            f.argtypes = [C.c_void_p] + a
            f.restype = None
        L.jo_spec_next_pow2.argtypes = [C.c_void_p, C.c_float]
        L.jo_spec_next_pow2.restype = C.c_size_t
        for n in ("spectrum_size", "memory_size", "feed_samples", "feed_blocks"):
            f = getattr(L, "jo_spec_" + n)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int
        L.jo_spec_samplerate.argtypes = [C.c_void_p]
        L.jo_spec_samplerate.restype = C.c_float
        L.jo_spec_process_block.argtypes = [C.c_void_p, f32p]
        L.jo_spec_get_mem.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
        _bind_pal(L, "jo_pal_")
        L.jo_pal_table.argtypes = [C.c_void_p, i32p, C.c_int]
        L.jo_pal_get_range.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 3
        L.jo_view_create.argtypes = [C.c_void_p, C.c_void_p]
        L.jo_view_create.restype = C.c_void_p
        L.jo_view_destroy.argtypes = [C.c_void_p]
        L.jo_view_set_running.argtypes = [C.c_void_p, C.c_int]
        L.jo_view_set_color_range.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.jo_view_force_recompute.argtypes = [C.c_void_p]
        L.jo_view_tick.argtypes = [C.c_void_p]
        L.jo_view_width.argtypes = [C.c_void_p]
        L.jo_view_height.argtypes = [C.c_void_p]
        L.jo_view_pixels.argtypes = [C.c_void_p]
        L.jo_view_pixels.restype = C.POINTER(C.c_uint32)
        L.jo_render_batch.argtypes = [C.POINTER(BatchCfg), f32p, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_void_p]
        L.jo_render_batch.restype = C.c_long
        L.jo_bench_batch.argtypes = [C.POINTER(BatchCfg), f32p, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long)]
        L.jo_bench_batch.restype = C.c_double
        L.jo_bench_stream.argtypes = [C.c_void_p, f32p, C.c_int]
        L.jo_bench_stream.restype = C.c_double
        _lib = L
    return _lib


def _bind_pal(L, pre):
    getattr(L, pre + "create").argtypes = [C.c_int, C.c_int]
    getattr(L, pre + "create").restype = C.c_void_p
    getattr(L, pre + "create_default").restype = C.c_void_p
    getattr(L, pre + "destroy").argtypes = [C.c_void_p]
    getattr(L, pre + "set_value_range").argtypes = [C.c_void_p, C.c_float, C.c_float]
    getattr(L, pre + "set_nr_of_colors").argtypes = [C.c_void_p, C.c_int]
    getattr(L, pre + "set_color_scheme").argtypes = [C.c_void_p, C.c_int]
    getattr(L, pre + "set_invert").argtypes = [C.c_void_p, C.c_int]
    getattr(L, pre + "get_rgb").argtypes = [C.c_void_p, C.c_float]
    getattr(L, pre + "get_rgb").restype = C.c_int
    getattr(L, pre + "get_value").argtypes = [C.c_void_p, C.c_int]
    getattr(L, pre + "get_value").restype = C.c_float
    getattr(L, pre + "lookup_many").argtypes = [C.c_void_p, f32p, C.c_int, i32p]


def ref():
    """Compiled reference sources (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        so = ORACLE_DIR / "_ref" / "libjade_ref.so"
        if not so.exists():
            return None
        lib()  # libjade_ref.so resolves jo_power_f32 (the FFT stand-in) from liboracle.so
        L = C.CDLL(str(so))
        _bind_pal(L, "jr_pal_")
        L.has_spec = hasattr(L, "jr_spec_create")
        if L.has_spec:
            _bind_ref_spec(L)
        _ref = L
    return _ref


def _bind_ref_spec(L):
    """The reference's real Spectrogram / SpectrogramComponent (oracle/ref_glue.cpp)."""
    L.jr_spec_create.restype = C.c_void_p
    for n, a in [("destroy", []), ("set_samplerate", [C.c_float]), ("set_channels", [C.c_size_t]),
                 ("set_fftsize", [C.c_size_t]), ("set_closest_fftsize_ms", [C.c_float]),
                 ("set_memory_time_s", [C.c_float]), ("set_feed_percent", [C.c_int]), ("set_pause", [C.c_int]),
                 ("set_window", [C.c_int]), ("set_mix_mode", [C.c_int])]:
        f = getattr(L, "jr_spec_" + n)
        f.argtypes = [C.c_void_p] + a
        f.restype = None
    L.jr_spec_next_pow2.argtypes = [C.c_void_p, C.c_float]
    L.jr_spec_next_pow2.restype = C.c_size_t
    for n in ("spectrum_size", "memory_size", "feed_samples", "feed_blocks"):
        f = getattr(L, "jr_spec_" + n)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
    L.jr_spec_samplerate.argtypes = [C.c_void_p]
    L.jr_spec_samplerate.restype = C.c_float
    L.jr_spec_window.argtypes = [C.c_void_p, f32p, C.c_int]
    L.jr_spec_process_block.argtypes = [C.c_void_p, f32p]
    L.jr_spec_get_mem.argtypes = [C.c_void_p, f32p, C.c_int, C.POINTER(C.c_int)]
    L.jr_view_create.argtypes = [C.c_void_p]
    L.jr_view_create.restype = C.c_void_p
    L.jr_view_destroy.argtypes = [C.c_void_p]
    L.jr_view_set_running.argtypes = [C.c_void_p, C.c_int]
    L.jr_view_set_color_range.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.jr_view_set_scheme.argtypes = [C.c_void_p, C.c_int]
    L.jr_view_force_recompute.argtypes = [C.c_void_p]
    L.jr_view_tick.argtypes = [C.c_void_p]
    L.jr_view_tick.restype = None
    L.jr_view_width.argtypes = [C.c_void_p]
    L.jr_view_height.argtypes = [C.c_void_p]
    L.jr_view_pixels.argtypes = [C.c_void_p]
    L.jr_view_pixels.restype = C.POINTER(C.c_uint32)
    if hasattr(L, "jr_view_paint"):
        L.jr_view_paint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_int), f32p,
                                    np.ctypeslib.ndpointer(np.int32, flags="C"), np.ctypeslib.ndpointer(np.uint32, flags="C"), C.c_int]
        L.jr_view_min_display_freq.argtypes = [C.c_void_p]
        L.jr_view_min_display_freq.restype = C.c_float
        L.jr_view_max_display_freq.argtypes = [C.c_void_p]
        L.jr_view_max_display_freq.restype = C.c_float
    L.jr_bench_batch.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                 f32p, C.c_long, C.c_int, C.c_int, C.POINTER(C.c_long)]
    L.jr_bench_batch.restype = C.c_double


def have_ref_spec():
    r = ref()
    return r is not None and r.has_spec


class Palette:
    """Same call surface over either the restated (oracle) or the compiled reference CColorPalette."""

    def __init__(self, n=None, scheme=None, use_ref=False):
        self.L = ref() if use_ref else lib()
        self.pre = "jr_pal_" if use_ref else "jo_pal_"
        if n is None:
            self.h = getattr(self.L, self.pre + "create_default")()
            self.n = 2
        else:
            self.h = getattr(self.L, self.pre + "create")(n, scheme)
            self.n = n

    def _c(self, name, *a):
        return getattr(self.L, self.pre + name)(self.h, *a)

    def set_value_range(self, a, b):
        self._c("set_value_range", a, b)

    def set_nr_of_colors(self, n):
        self.n = n
        self._c("set_nr_of_colors", n)

    def set_color_scheme(self, s):
        self._c("set_color_scheme", s)

    def set_invert(self, on):
        self._c("set_invert", int(on))

    def get_rgb(self, v):
        return self._c("get_rgb", float(v))

    def get_value(self, c):
        return self._c("get_value", int(c))

    def lookup(self, v):
        v = np.ascontiguousarray(v, np.float32)
        out = np.empty(v.size, np.int32)
        self._c("lookup_many", v.reshape(-1), v.size, out)
        return out.reshape(v.shape)

    def table(self):
        """Palette table recovered through the public lookup only (works for the reference too)."""
        n = self.n
        self.set_value_range(0.0, float(n))
        return self.lookup(np.arange(n, dtype=np.float32) + 0.5)

    def __del__(self):
        try:
            self._c("destroy")
        except Exception:
            pass


class Spec:
    """Restated Spectrogram (oracle), or with use_ref=True the reference's REAL class compiled in oracle/_ref."""

    def __init__(self, use_ref=False):
        self.L = ref() if use_ref else lib()
        self.pre = "jr_spec_" if use_ref else "jo_spec_"
        self.h = getattr(self.L, self.pre + "create")()

    def __getattr__(self, name):
        f = getattr(self.L, self.pre + name)
        return lambda *a: f(self.h, *a)

    def process(self, planar):
        planar = np.ascontiguousarray(planar, np.float32)
        return getattr(self.L, self.pre + "process_block")(self.h, planar.reshape(-1))

    def get_mem(self, mem):
        pos = C.c_int(0)
        r = getattr(self.L, self.pre + "get_mem")(self.h, mem.reshape(-1), mem.shape[0], C.byref(pos))
        return r, pos.value

    def __del__(self):
        try:
            getattr(self.L, self.pre + "destroy")(self.h)
        except Exception:
            pass


class View:
    """SpectrogramComponent::timerCallback image assembly: restated (oracle) or the reference's real one."""

    def __init__(self, spec, pal=None, use_ref=False):
        self.use_ref = use_ref
        self.spec = spec
        self.pal = pal
        if use_ref:
            self.L = ref()
            self.pre = "jr_view_"
            self.h = self.L.jr_view_create(spec.h)
        else:
            self.L = lib()
            self.pre = "jo_view_"
            self.h = self.L.jo_view_create(spec.h, pal.h)

    def _c(self, name, *a):
        return getattr(self.L, self.pre + name)(self.h, *a)

    def tick(self):
        return self._c("tick")

    def set_running(self, on):
        self._c("set_running", int(on))

    def set_color_range(self, mn, mx):
        self._c("set_color_range", float(mn), float(mx))
        if not self.use_ref:
            self._c("force_recompute")

    def force_recompute(self):
        self._c("force_recompute")

    def set_scheme(self, idx):
        assert self.use_ref
        self._c("set_scheme", int(idx))

    def paint(self, width, height, scale, min_hz, max_hz):
        """The reference's REAL SpectrogramComponent::paint (Spectrogram.cpp:432-545) on the recording Graphics stub:
        dict(crop=(hStart, heightInterval), freq_val, freq_y, color_val, color_y, colorbar, min_hz, max_hz)."""
        assert self.use_ref
        crop = (C.c_int * 2)()
        val = np.zeros(22, np.float32)
        y = np.zeros(22, np.int32)
        cb = np.zeros(4096, np.uint32)
        n = self._c("paint", int(width), int(height), float(scale), float(np.log(np.float32(min_hz))),
                    float(np.log(np.float32(max_hz))), crop, val, y, cb, cb.size)
        assert n > 0, n
        return dict(crop=(crop[0], crop[1]), freq_val=val[:11].copy(), freq_y=y[:11].copy(), color_val=val[11:].copy(),
                    color_y=y[11:].copy(), colorbar=cb[:n].copy(), min_hz=self._c("min_display_freq"), max_hz=self._c("max_display_freq"))

    def image(self):
        h, w = self._c("height"), self._c("width")
        return np.ctypeslib.as_array(self._c("pixels"), shape=(h, w)).copy()

    def __del__(self):
        try:
            self._c("destroy")
        except Exception:
            pass


def window(kind, n):
    out = np.empty(n, np.float32)
    assert lib().jo_window(WIN[kind] if isinstance(kind, str) else kind, n, out) == 0
    return out


def power_f32(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.size // 2 + 1, np.float32)
    assert lib().jo_power_f32(x, x.size, out) == 0
    return out


def power_f64(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.size // 2 + 1, np.float64)
    assert lib().jo_power_f64(x, x.size, out) == 0
    return out


def render_batch(samples, *, fs=48000.0, fft_size=2048, hop=512, window="hann", mix="absmean", scheme="jade",
                 ncolors=256, invert=False, min_db=-50.0, max_db=50.0, first_col=0, ncols=None, double_fft=False,
                 want_db=True, want_pix=True):
    """samples [channels][nsamples] -> (db [ncols][B] float32, pix [ncols][B] uint32 ARGB, row r <-> bin B-1-r)."""
    samples = np.ascontiguousarray(np.atleast_2d(samples), np.float32)
    ch, ns = samples.shape
    B = fft_size // 2 + 1
    if ncols is None:
        ncols = ns // hop + 1
    cfg = BatchCfg(fs, fft_size, hop, WIN[window], ch, MIX[mix], PAL[scheme], ncolors, int(invert), min_db, max_db,
                   int(double_fft))
    db = np.empty((ncols, B), np.float32) if want_db else None
    pix = np.empty((ncols, B), np.uint32) if want_pix else None
    r = lib().jo_render_batch(C.byref(cfg), samples.reshape(-1), ns, first_col, ncols,
                              db.ctypes.data if want_db else None, pix.ctypes.data if want_pix else None)
    assert r == ncols, r
    return db, pix
