"""Oracle pins for the STFT core: windows against the committed goldens, known-answer tests that pin
the unnormalised forward DFT convention, the restated Spectrogram class (framing, perc10, ring, getMem, pause) and the
product's host-side window / row-map tables (host-only entry points of libjade_gpu.so)."""
import ctypes as C
import hashlib
import pathlib
import sys

import numpy as np
import pytest

import oracle_lib as O
import signals

GOLD_W = np.load(pathlib.Path(__file__).parent / "golden" / "windows.npz")
GOLD_P = np.load(pathlib.Path(__file__).parent / "golden" / "pipeline.npz")


# ---------------------------------------------------------------- windows
@pytest.mark.parametrize("name", list(O.WIN))
def test_window_tables_bit_exact(name):
    from jadespectrogram_b200 import _capi
    lib = _capi.load()
    for n in (64, 2048):
        w = O.window(name, n)
        assert np.array_equal(w, GOLD_W[f"{name}_{n}"])
        p = np.empty(n, np.float32)
        assert lib.jade_window_build(O.WIN[name], n, p.ctypes.data) == 0
        assert np.array_equal(p, w), f"product window {name} N={n} differs from the oracle"
    big = np.empty(65536, np.float32)
    lib.jade_window_build(O.WIN[name], 65536, big.ctypes.data)
    for tab in (O.window(name, 65536), big):
        assert np.array_equal(np.frombuffer(hashlib.sha256(tab.tobytes()).digest(), np.uint8), GOLD_W[f"{name}_65536_sha256"])


def test_window_properties():
    for name in O.WIN:
        w = O.window(name, 4096).astype(np.float64)
        assert abs(np.sqrt(np.mean(w * w)) - 1.0) < 2e-5  # unit RMS (float32 accumulation error)
    hp = O.window("hannpoisson", 1024)
    assert (hp[513:] == 0).all() and (hp[1:512] > 0).all()  # size_t wrap quirk (Spectrogram.cpp:280)
    assert (O.window("rect", 64) == 1.0).all()
    # Hann is periodic: w[k] == w[N-k]
    h = O.window("hann", 2048)
    assert np.allclose(h[1:], h[1:][::-1], atol=1e-6)


# ---------------------------------------------------------------- FFT stand-in known answers (unnormalised DFT)
def test_fft_known_answers():
    N = 2048
    n = np.arange(N)
    w = O.window("hann", N).astype(np.float64)
    for k0, A in ((64, 0.5), (300, 1.0), (1000, 0.25)):
        x = (A * np.cos(2 * np.pi * k0 * n / N) * w).astype(np.float32)
        p = O.power_f32(x)
        assert p[k0] == pytest.approx((A * w.sum() / 2) ** 2, rel=2e-5)
    # DC
    x = (0.3 * w).astype(np.float32)
    assert O.power_f32(x)[0] == pytest.approx((0.3 * w.sum()) ** 2, rel=2e-5)
    # unit impulse -> flat spectrum w[n0]^2
    x = np.zeros(N, np.float32)
    x[77] = w[77]
    assert np.allclose(O.power_f32(x), w[77] ** 2, rtol=1e-4)
    # Parseval with the one-sided weights
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N).astype(np.float32)
    p = O.power_f64(x)
    assert p[0] + p[-1] + 2 * p[1:-1].sum() == pytest.approx(N * float((x.astype(np.float64) ** 2).sum()), rel=1e-9)
    # against numpy in float64 and the float32 error bound used by the parity tests
    t = np.abs(np.fft.rfft(x.astype(np.float64))) ** 2
    assert np.allclose(p, t, rtol=1e-9, atol=1e-9)
    err = np.abs(np.sqrt(O.power_f32(x)) - np.sqrt(t))
    assert err.max() < 8 * np.finfo(np.float32).eps * np.sqrt(np.log2(N) * t.sum() / N)
    assert O.lib().jo_power_f32(np.zeros(12, np.float32), 12, np.zeros(7, np.float32)) != 0  # not a power of two


def test_db_floor():
    assert O.lib().jo_db(0.0) == np.float32(10.0 * np.log10(np.float64(np.float32(1e-11))))
    assert abs(O.lib().jo_db(0.0) + 110.0) < 1e-5
    assert O.lib().jo_db(1.0) == pytest.approx(0.0, abs=1e-6)


# ---------------------------------------------------------------- restated Spectrogram
def _spec(N=1024, feed="p50", ch=1, fs=48000.0, mem=1.0):
    s = O.Spec()
    s.set_channels(ch)
    s.set_samplerate(fs)
    s.set_memory_time_s(mem)
    s.set_fftsize(N)
    s.set_feed_percent(O.FEED[feed])
    return s


def test_buildmem_geometry():
    s = O.Spec()  # constructor defaults (Spectrogram.cpp:16-24)
    assert (s.spectrum_size(), s.feed_samples(), s.feed_blocks(), s.memory_size()) == (513, 1024, 1, 47)
    s = _spec(2048, "p50", 2, 48000.0, 10.0)  # the plugin's prepareToPlay (PluginProcessor.cpp:102-114)
    assert (s.spectrum_size(), s.feed_samples(), s.feed_blocks(), s.memory_size()) == (1025, 1024, 2, 469)
    s.set_feed_percent(O.FEED["p10"])
    assert (s.feed_samples(), s.feed_blocks()) == (205, 10)
    assert s.next_pow2(20.0) == 1024 and s.next_pow2(40.0) == 2048  # 960 -> 1024, 1920 -> 2048


def test_first_getmem_reports_everything_new_and_ring_is_minus_120():
    s = _spec(1024, "p50", 1)
    W, B = s.memory_size(), s.spectrum_size()
    mem = np.zeros((W, B), np.float32)
    newv, pos = s.get_mem(mem)
    assert newv == 1215752192 and pos == 0  # int(100000000000)
    assert (mem == -120.0).all()
    assert s.get_mem(np.zeros((W + 1, B), np.float32))[0] == -1
    assert s.get_mem(mem)[0] == 0


def test_column_geometry_and_batch_equivalence():
    """column j analyses x[j*hop - N, j*hop): the streaming class and the batch helper agree bit for bit."""
    N, hop = 1024, 512
    s = _spec(N, "p50", 1)
    W, B = s.memory_size(), s.spectrum_size()
    x = signals.streams(1, 1, N * 6, 48000.0)[0]
    mem = np.zeros((W, B), np.float32)
    s.get_mem(mem)
    for b in range(6):
        s.process(x[:, b * N:(b + 1) * N])
    newv, pos = s.get_mem(mem)
    assert (newv, pos) == (12, 12)
    db, _ = O.render_batch(x, fft_size=N, hop=hop, ncols=12)
    assert np.array_equal(mem[:12], db)
    assert np.allclose(db[0], -110.0, atol=1e-5)  # column 0 sees only the zero pre-roll


def test_perc10_non_uniform_hop():
    N = 2048
    s = _spec(N, "p10", 1)
    x = signals.streams(1, 1, N * 2, 48000.0)[0]
    W, B = s.memory_size(), s.spectrum_size()
    mem = np.zeros((W, B), np.float32)
    s.get_mem(mem)
    s.process(x[:, :N])
    s.process(x[:, N:])
    newv, pos = s.get_mem(mem)
    assert newv == 20
    # sub-frame bb of block b starts at b*N + 205*bb - N: check two of them against a direct computation
    w = O.window("hann", N)
    P = np.concatenate([np.zeros(N, np.float32), x[0]])
    for col, start in ((3, 3 * 205), (10, N), (19, N + 9 * 205)):
        frame = (P[start:start + N] * w).astype(np.float32)
        ref = np.array([O.lib().jo_db(v) for v in O.power_f32(frame)], np.float32)
        assert np.array_equal(mem[col], ref)


def test_ring_wrap_and_pause():
    N = 512
    s = _spec(N, "p100", 1, 48000.0, 0.05)  # tiny ring
    W, B = s.memory_size(), s.spectrum_size()
    assert W == 5
    x = signals.streams(1, 1, N * 9, 48000.0)[0]
    mem = np.zeros((W, B), np.float32)
    s.get_mem(mem)
    for b in range(3):
        s.process(x[:, b * N:(b + 1) * N])
    assert s.get_mem(mem) == (3, 3)
    s.set_pause(1)
    s.process(x[:, 3 * N:4 * N])  # FFT still runs, ring untouched (Spectrogram.cpp:111-118)
    assert s.get_mem(mem) == (0, 3)
    s.set_pause(0)
    for b in range(4, 8):
        s.process(x[:, b * N:(b + 1) * N])
    newv, pos = s.get_mem(mem)
    assert (newv, pos) == (4, 2)  # wrapped: slots 3,4,0,1
    db, _ = O.render_batch(x, fft_size=N, hop=N, ncols=9)
    assert np.array_equal(mem[3], db[4]) and np.array_equal(mem[1], db[7]) and np.array_equal(mem[2], db[2])


def test_mix_modes():
    N = 256
    x = signals.streams(1, 2, N * 3, 48000.0)[0]
    x[1] *= 3.0
    res = {m: O.render_batch(x, fft_size=N, hop=N, mix=m, ncols=3)[0] for m in O.MIX}
    left = O.render_batch(x[:1], fft_size=N, hop=N, ncols=3)[0]
    right = O.render_batch(x[1:], fft_size=N, hop=N, ncols=3)[0]
    assert np.array_equal(res["left"], left) and np.array_equal(res["right"], right)
    assert np.array_equal(res["max"], np.maximum(left, right)) and np.array_equal(res["min"], np.minimum(left, right))
    assert ((res["absmean"] <= res["max"] + 1e-4) & (res["absmean"] >= res["min"] - 1e-4)).all()


def test_pipeline_goldens():
    for name in ("n64_mono", "n256_stereo", "n1024_cfg1", "n2048_cfg2", "n512_max3"):
        N, hop, ch, win, mix = (int(v) for v in GOLD_P[name + "_cfg"])
        winname = [k for k, v in O.WIN.items() if v == win][0]
        mixname = [k for k, v in O.MIX.items() if v == mix][0]
        db, pix = O.render_batch(GOLD_P[name + "_x"], fft_size=N, hop=hop, window=winname, mix=mixname,
                                 ncols=GOLD_P[name + "_db"].shape[0])
        assert np.array_equal(db, GOLD_P[name + "_db"]) and np.array_equal(pix, GOLD_P[name + "_pix"])


# ---------------------------------------------------------------- image assembly restatement (Spectrogram.cpp:590-724)
def test_view_scroll_and_fix_modes():
    N = 256
    s = _spec(N, "p100", 1, 48000.0, 0.05)
    W, B = s.memory_size(), s.spectrum_size()
    pal = O.Palette(256, O.PAL["jade"])
    L = O.lib()
    v = L.jo_view_create(s.h, pal.h)
    x = signals.streams(1, 1, N * 12, 48000.0)[0]

    def image():
        return np.ctypeslib.as_array(L.jo_view_pixels(v), shape=(L.jo_view_height(v), L.jo_view_width(v))).copy()

    L.jo_view_tick(v)  # first tick: full redraw of the -120 dB ring
    pal.set_value_range(-50, 50)
    floor_col = np.uint32(pal.get_rgb(-120.0)) | np.uint32(0xFF000000)
    assert (image() == floor_col).all()
    for b in range(3):
        s.process(x[:, b * N:(b + 1) * N])
    assert L.jo_view_tick(v) == 3
    img = image()
    db, pix = O.render_batch(x, fft_size=N, hop=N, ncols=12)
    # scroll mode: newest column is the right-most, rows are flipped bins (Spectrogram.cpp:642,665-682)
    assert np.array_equal(img[:, W - 1], pix[2]) and np.array_equal(img[:, W - 3], pix[0])
    assert (img[:, :W - 3] == floor_col).all()
    # "Fix" mode writes at the ring position and draws a red cursor
    L.jo_view_set_running(v, 0)
    s.process(x[:, 3 * N:4 * N])
    assert L.jo_view_tick(v) == 1
    img = image()
    assert np.array_equal(img[:, 3], pix[3])
    assert (img[:, 4 % W] == 0xFFFF0000).all()
    L.jo_view_destroy(v)


# ---------------------------------------------------------------- product host tables (no GPU needed)
def test_linear_crop_matches_paint_maths():
    from jadespectrogram_b200 import _capi
    lib = _capi.load()
    lo, hi = C.c_int(), C.c_int()
    fs, H = 48000.0, 1025
    for fmin, fmax in ((1.0, 20000.0), (100.0, 5000.0), (30000.0, 40000.0), (5000.0, 100.0), (0.0, 24000.0)):
        assert lib.jade_linear_crop(fs, H, fmin, fmax, C.byref(lo), C.byref(hi)) == 0
        a, b = np.float32(fmin), np.float32(fmax)
        if a >= fs * 0.5:
            a = np.float32(0.9 * fs * 0.5)
        if b >= fs * 0.5:
            b = np.float32(fs * 0.5)
        if a >= b:
            a = np.float32(0.9 * float(b))
        end = int(2.0 * float(b) / fs * H + 0.5)
        interval = int(2.0 * float(b) / fs * H - 2.0 * float(a) / fs * H + 0.5)
        # the engine never returns an empty crop (the reference would draw zero rows): at least one row
        assert hi.value == min(end, H)
        assert hi.value - lo.value == max(1, interval - max(0, end - H)) or lo.value == 0
        assert 0 <= lo.value < hi.value <= H


def test_log_rows_cover_the_band_monotonically():
    from jadespectrogram_b200 import _capi
    lib = _capi.load()
    R, N, fs = 1080, 65536, 192000.0
    lo, hi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    assert lib.jade_log_rows(fs, N, R, 20.0, 96000.0, lo.ctypes.data, hi.ctypes.data) == 0
    assert (hi > lo).all() and (lo >= 0).all() and (hi <= N // 2 + 1).all()
    assert (np.diff(lo) >= 0).all() and (np.diff(hi) >= 0).all()
    assert lo[0] == round(20.0 * (96000 / 20) ** (0.5 / R) / (fs / N)) or lo[0] in (6, 7)
    assert hi[-1] == N // 2 + 1 or hi[-1] == N // 2
    wide = hi - lo > 1
    assert (lo[1:][wide[1:]] == hi[:-1][wide[1:]]).all()  # once bands are wider than a bin they tile without gaps


# ---------------------------------------------------------------- goldens produced by the reference's REAL class
GOLD_R = np.load(pathlib.Path(__file__).parent / "golden" / "spectrogram_ref.npz")


def test_oracle_matches_reference_class_goldens():
    """tests/golden/spectrogram_ref.npz holds outputs of the reference's own Spectrogram / timerCallback code compiled in
    place (tools/gen_golden.py: reference_class); the restated oracle must reproduce them bit for bit."""
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent / "tools"))
    from gen_golden import REF_CASES, ref_case_inputs
    for name in O.WIN:
        for n in (64, 1024):
            assert np.array_equal(O.window(name, n).view(np.uint32), GOLD_R[f"window_{name}_{n}"].view(np.uint32)), (name, n)
    for name, N, feed, ch, win, nblocks in REF_CASES:
        cfg = GOLD_R[name + "_cfg"]
        s = O.Spec()
        s.set_samplerate(48000.0)
        s.set_memory_time_s(0.1)
        s.set_channels(ch)
        s.set_fftsize(N)
        s.set_feed_percent(O.FEED[feed])
        s.set_window(O.WIN[win])
        W, B = s.memory_size(), s.spectrum_size()
        assert (W, B) == (int(cfg[5]), int(cfg[6]))
        pal = O.Palette(256, O.PAL["jade"])
        view = O.View(s, pal)
        view.tick()
        x = ref_case_inputs(N, ch, nblocks)
        for b in range(nblocks):
            s.process(x[:, b * N:(b + 1) * N])
        view.tick()
        assert np.array_equal(view.image(), GOLD_R[name + "_image"]), name
        for b in range(nblocks):
            s.process(x[:, b * N:(b + 1) * N])
        mem = np.zeros((W, B), np.float32)
        newvals, pos = s.get_mem(mem)
        assert (newvals, pos) == (int(cfg[7]), int(cfg[8])), name
        assert np.array_equal(mem.view(np.uint32), GOLD_R[name + "_ring_db"].view(np.uint32)), name
