"""GPU tests of the drop-in C++ classes (include/Spectrogram.h, include/CColorpalette.h) driven exactly like the plugin
drives the reference (PluginProcessor.cpp:102-114,148; Spectrogram.cpp:590-724), against the restated oracle class AND against
the reference's REAL class (oracle/_ref/libjade_ref.so: Spectrogram.cpp compiled in place; it travels to the GPU box prebuilt)."""
import ctypes as C
import pathlib

import numpy as np
import pytest

import oracle_lib as O
import parity
import signals

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


class Dropin:
    def __init__(self):
        L = C.CDLL(str(ROOT / "jadespectrogram_b200" / "libjade_dropin_shim.so"))
        L.jd_spec_create.restype = C.c_void_p
        L.jd_spec_error.restype = C.c_char_p
        L.jd_spec_samplerate.restype = C.c_float
        L.jd_spec_next_pow2.restype = C.c_size_t
        for name, args in [("destroy", []), ("ok", []), ("error", []), ("set_samplerate", [C.c_float]), ("set_channels", [C.c_size_t]),
                           ("set_fftsize", [C.c_size_t]), ("set_closest_fftsize_ms", [C.c_float]), ("set_memory_time_s", [C.c_float]),
                           ("set_feed_percent", [C.c_int]), ("set_pause", [C.c_int]), ("set_window", [C.c_int]),
                           ("set_mix_mode", [C.c_int]), ("next_pow2", [C.c_float]), ("spectrum_size", []), ("memory_size", []),
                           ("samplerate", []), ("process_block", [C.c_void_p, C.c_int, C.c_int]), ("prepare", [C.c_int, C.c_int]),
                           ("process_audio", [C.c_void_p, C.c_int, C.c_int]), ("get_mem", [C.c_void_p, C.c_int, C.POINTER(C.c_int)])]:
            getattr(L, "jd_spec_" + name).argtypes = [C.c_void_p] + args
        self.L = L
        self.h = L.jd_spec_create()
        assert L.jd_spec_ok(self.h) == 1, L.jd_spec_error(self.h)

    def __getattr__(self, name):
        f = getattr(self.L, "jd_spec_" + name)
        return lambda *a: f(self.h, *a)

    def process(self, planar):
        planar = np.ascontiguousarray(planar, np.float32)
        return self.L.jd_spec_process_block(self.h, planar.ctypes.data, planar.shape[0], planar.shape[1])

    def audio(self, planar):
        planar = np.ascontiguousarray(planar, np.float32)
        return self.L.jd_spec_process_audio(self.h, planar.ctypes.data, planar.shape[0], planar.shape[1])

    def get_mem(self, mem):
        pos = C.c_int(0)
        r = self.L.jd_spec_get_mem(self.h, mem.ctypes.data, mem.shape[0], C.byref(pos))
        return r, pos.value

    def close(self):
        self.L.jd_spec_destroy(self.h)


def _prepare(obj, fs, ch, N, feed, mem_s=1.0):
    """The plugin's prepareToPlay sequence (PluginProcessor.cpp:108-112)."""
    obj.set_channels(ch)
    obj.set_samplerate(fs)
    obj.set_memory_time_s(mem_s)
    obj.set_fftsize(N)
    obj.set_feed_percent(O.FEED[feed])


ARMS = ["oracle", pytest.param("reference", marks=pytest.mark.skipif(not O.have_ref_spec(), reason="oracle/_ref not built"))]


@pytest.mark.parametrize("arm", ARMS)
@pytest.mark.parametrize("N,feed,ch", [(1024, "p50", 1), (2048, "p50", 2), (2048, "p25", 2), (2048, "p10", 1), (512, "p100", 2), (8192, "p25", 1)])
def test_dropin_class_follows_reference_class(N, feed, ch, arm):
    fs = 48000.0
    d, o = Dropin(), O.Spec(use_ref=(arm == "reference"))
    _prepare(d, fs, ch, N, feed)
    _prepare(o, fs, ch, N, feed)
    assert (d.spectrum_size(), d.memory_size(), d.samplerate()) == (o.spectrum_size(), o.memory_size(), o.samplerate())
    assert d.next_pow2(20.0) == o.next_pow2(20.0)
    W, B = o.memory_size(), o.spectrum_size()
    md, mo = np.zeros((W, B), np.float32), np.zeros((W, B), np.float32)
    assert d.get_mem(md) == o.get_mem(mo) == (1215752192, 0)
    assert np.array_equal(md, mo) and (md == -120.0).all()
    assert d.get_mem(np.zeros((W + 1, B), np.float32))[0] == -1
    nblocks = 2 * W // o.feed_blocks() + 3  # wraps the ring
    x = signals.streams(1, ch, N * nblocks, fs)[0]
    calls = 0
    for b in range(nblocks):
        blk = x[:, b * N:(b + 1) * N]
        assert d.process(blk) == 0
        o.process(blk)
        if b % 3 == 2 or b == nblocks - 1:  # the GUI timer fires every few blocks
            rd, ro = d.get_mem(md), o.get_mem(mo)
            assert rd == ro, (b, rd, ro)
            parity.check_db(md, mo, N, f"block {b}")
            calls += 1
    assert calls > 3
    d.close()


@pytest.mark.parametrize("arm", ARMS)
def test_pause_window_change_and_reblocker(arm):
    fs, N, ch = 48000.0, 1024, 2
    d, o = Dropin(), O.Spec(use_ref=(arm == "reference"))
    _prepare(d, fs, ch, N, "p50")
    _prepare(o, fs, ch, N, "p50")
    W, B = o.memory_size(), o.spectrum_size()
    md, mo = np.zeros((W, B), np.float32), np.zeros((W, B), np.float32)
    d.get_mem(md), o.get_mem(mo)
    x = signals.streams(1, ch, N * 10, fs)[0]
    # host blocks of 480 samples through the re-blocker == N-sample blocks through processSynchronBlock
    d.prepare(ch, 480)
    pos = 0
    while pos + 480 <= N * 4:
        assert d.audio(x[:, pos:pos + 480]) == 0
        pos += 480
    for b in range(pos // N):
        o.process(x[:, b * N:(b + 1) * N])
    assert d.get_mem(md) == o.get_mem(mo)
    parity.check_db(md, mo, N, "reblocked")
    # finish the partially filled block so both sit on a block boundary again
    rest = (pos // N + 1) * N - pos
    d.audio(x[:, pos:pos + rest])
    o.process(x[:, (pos // N) * N:(pos // N + 1) * N])
    nxt = pos // N + 1
    # pause: counters and ring stand still (Spectrogram.cpp:111-118)
    d.set_pause(1), o.set_pause(1)
    d.process(x[:, nxt * N:(nxt + 1) * N]), o.process(x[:, nxt * N:(nxt + 1) * N])
    d.set_pause(0), o.set_pause(0)
    # window change without buffer reset (Spectrogram.h:123)
    d.set_window(O.WIN["blackmanharris"]), o.set_window(O.WIN["blackmanharris"])
    for b in range(nxt + 1, nxt + 4):
        d.process(x[:, b * N:(b + 1) * N]), o.process(x[:, b * N:(b + 1) * N])
    rd, ro = d.get_mem(md), o.get_mem(mo)
    assert rd == ro
    parity.check_db(md, mo, N, "after pause + window change")
    d.close()


def test_fewer_input_channels_is_an_error_not_ub():
    d = Dropin()
    _prepare(d, 48000.0, 2, 1024, "p50")
    assert d.process(np.zeros((1, 1024), np.float32)) == -1  # the reference reads out of bounds here
    assert b"channels" in d.error()
    d.close()
