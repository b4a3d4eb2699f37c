"""The restated oracle against the reference's REAL code, executed here.

oracle/_ref/libjade_ref.so compiles /root/reference/Spectrogram.cpp and CColorpalette.cpp where they lie, against the
inert JUCE/TGM stubs of oracle/shim/ (oracle/Makefile).  Everything the reference itself contains on the hot path --
buildmem geometry, the six windows, framing / staging, the channel-mix switch, dB, the ring, getMem, and the
timerCallback pixel loops -- therefore runs as written by its author and must agree BIT FOR BIT with the restatement in
oracle/jade_oracle.cpp.  The one thing that is not the reference's is the FFT (`spectrum`, absent TGM library): both
sides use the oracle's float32 stand-in, so FFT parity stays unpinned.

Skipped when oracle/_ref was built without the Spectrogram glue (never the case for the committed recipe).
"""
import numpy as np
import pytest

import oracle_lib as O
import signals

# also_gpu: CPU-only, but selected by the driver's `-m gpu` run too (tests/conftest.py), so that its record holds the whole pin
pytestmark = [pytest.mark.skipif(not O.have_ref_spec(), reason="oracle/_ref/libjade_ref.so (compiled reference) not built"),
              pytest.mark.also_gpu]

FS = 48000.0


def _pair(N, feed, ch, fs=FS, mem_s=0.25, window="hann", mix=None):
    out = []
    for use_ref in (False, True):
        s = O.Spec(use_ref=use_ref)
        # the plugin's order (PluginProcessor.cpp:108-112)
        s.set_samplerate(fs)
        s.set_memory_time_s(mem_s)
        s.set_channels(ch)
        s.set_fftsize(N)
        s.set_feed_percent(O.FEED[feed])
        s.set_window(O.WIN[window])
        if mix is not None:
            s.set_mix_mode(O.MIX[mix])
        out.append(s)
    return out


@pytest.mark.parametrize("window", list(O.WIN))
@pytest.mark.parametrize("N", [64, 512, 2048, 16384, 65536])
def test_window_tables(window, N):
    r = O.Spec(use_ref=True)
    r.set_fftsize(N)
    r.set_window(O.WIN[window])
    w = np.empty(N, np.float32)
    assert O.ref().jr_spec_window(r.h, w, N) == 0
    assert np.array_equal(w.view(np.uint32), O.window(window, N).view(np.uint32))


@pytest.mark.parametrize("feed", list(O.FEED))
@pytest.mark.parametrize("N,ch", [(256, 1), (1024, 2), (2048, 2)])
def test_geometry_and_columns(feed, N, ch):
    o, r = _pair(N, feed, ch)
    for name in ("spectrum_size", "memory_size", "feed_samples", "feed_blocks"):
        assert getattr(o, name)() == getattr(r, name)(), name
    W, B = o.memory_size(), o.spectrum_size()
    x = signals.streams(1, ch, N * 9, FS, kind="mix", seed=N + ch)[0]
    mo = np.zeros((W, B), np.float32)
    mr = np.zeros((W, B), np.float32)
    # the first getMem reports "everything new" (int(100000000000), Spectrogram.cpp:18,168) and hands out the -120 ring
    assert o.get_mem(mo) == r.get_mem(mr)
    assert np.array_equal(mo, mr) and (mr == -120.0).all()
    for b in range(9):
        blk = x[:, b * N:(b + 1) * N]
        assert o.process(blk) == r.process(blk) == 0
        if b % 2 == 1 or b == 8:
            assert o.get_mem(mo) == r.get_mem(mr)
            assert np.array_equal(mo.view(np.uint32), mr.view(np.uint32)), f"ring differs after block {b}"


@pytest.mark.parametrize("mix", list(O.MIX))
def test_mix_modes(mix):
    o, r = _pair(512, "p50", 2, mix=mix)
    W, B = o.memory_size(), o.spectrum_size()
    x = signals.streams(1, 2, 512 * 4, FS, kind="mix", seed=3)[0]
    x[1] *= 0.25
    mo, mr = np.zeros((W, B), np.float32), np.zeros((W, B), np.float32)
    for b in range(4):
        o.process(x[:, b * 512:(b + 1) * 512])
        r.process(x[:, b * 512:(b + 1) * 512])
    assert o.get_mem(mo) == r.get_mem(mr)
    assert np.array_equal(mo.view(np.uint32), mr.view(np.uint32))


def test_pause_ring_wrap_and_setwindow_midstream():
    o, r = _pair(256, "p25", 1, mem_s=0.02)  # a tiny ring: wraps several times
    W, B = o.memory_size(), o.spectrum_size()
    x = signals.streams(1, 1, 256 * 16, FS, kind="mix", seed=11)[0]
    mo, mr = np.zeros((W, B), np.float32), np.zeros((W, B), np.float32)
    o.get_mem(mo), r.get_mem(mr)
    for b in range(16):
        if b == 5:
            o.set_pause(1), r.set_pause(1)
        if b == 8:
            o.set_pause(0), r.set_pause(0)
        if b == 10:
            o.set_window(O.WIN["flattop"]), r.set_window(O.WIN["flattop"])
        o.process(x[:, b * 256:(b + 1) * 256])
        r.process(x[:, b * 256:(b + 1) * 256])
        if b in (2, 6, 9, 15):
            assert o.get_mem(mo) == r.get_mem(mr)
            assert np.array_equal(mo.view(np.uint32), mr.view(np.uint32)), b
    assert o.get_mem(np.zeros((W + 1, B), np.float32))[0] == r.get_mem(np.zeros((W + 1, B), np.float32))[0] == -1


def test_next_pow2_and_ms_setter():
    o, r = O.Spec(), O.Spec(use_ref=True)
    for fs in (44100.0, 48000.0, 96000.0):
        o.set_samplerate(fs), r.set_samplerate(fs)
        for ms in (1.0, 5.0, 10.7, 21.3, 42.7, 100.0):
            assert o.next_pow2(ms) == r.next_pow2(ms)
            o.set_closest_fftsize_ms(ms), r.set_closest_fftsize_ms(ms)
            assert o.spectrum_size() == r.spectrum_size() and o.memory_size() == r.memory_size()


@pytest.mark.parametrize("running", [1, 0])
def test_timer_callback_pixel_loops(running):
    """SpectrogramComponent::timerCallback as written (Spectrogram.cpp:590-731) vs the restated view, pixel for pixel."""
    N = 256
    o, r = _pair(N, "p50", 2, mem_s=0.1)
    pal = O.Palette(256, O.PAL["jade"])
    vo, vr = O.View(o, pal), O.View(r, use_ref=True)
    vo.set_running(running), vr.set_running(running)
    x = signals.streams(1, 2, N * 14, FS, kind="mix", seed=5)[0]
    vo.tick(), vr.tick()  # full redraw of the -120 ring
    assert np.array_equal(vo.image(), vr.image())
    W = o.memory_size()
    done = 0
    for nblk in (1, 3, 2, 1):
        for _ in range(nblk):
            blk = x[:, done * N:(done + 1) * N]
            o.process(blk), r.process(blk)
            done += 1
        vo.tick(), vr.tick()
        io, ir = vo.image(), vr.image()
        if not running:
            # The reference draws its red cursor at pos+dd and only wraps the == W case (Spectrogram.cpp:713-716), i.e. it
            # writes out of bounds when pos+dd > W; the restatement wraps.  Compare away from that corner.
            pos = (done * 2) % W
            if pos + 4 > W:
                continue
        assert np.array_equal(io, ir), (running, done)
    # range change -> m_recomputeAll (Spectrogram.cpp:370,379,623-657)
    vo.set_color_range(-80.0, 10.0), vr.set_color_range(-80.0, 10.0)
    vo.tick(), vr.tick()
    assert np.array_equal(vo.image(), vr.image())


def test_reference_bench_entry_point_counts_frames():
    import ctypes as C
    x = signals.streams(1, 2, 2048 * 6, FS, kind="mix")[0]
    frames = C.c_long(0)
    fps = O.ref().jr_bench_batch(FS, 2048, O.FEED["p25"], O.WIN["hann"], 2, O.PAL["jade"], 256, -50.0, 50.0, x.reshape(-1),
                                 x.shape[1], 4, 2, C.byref(frames))
    assert frames.value == 4 * 6 * 4 and fps > 0
