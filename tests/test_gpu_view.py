"""GPU test of the display-image assembly (jade_view_*, include/jade_gpu.h) against the restated
SpectrogramComponent::timerCallback of the oracle (itself pinned bit for bit against the reference's real one,
tests/test_oracle_vs_reference.py) and against the real one directly: scroll mode, fixed mode with the red cursor, full
redraws after a range change."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import signals
from jadespectrogram_b200 import Engine

pytestmark = pytest.mark.gpu
RED = 0xFFFF0000


class View:
    def __init__(self, eng):
        self.eng, self.lib = eng, eng.lib
        self.h = C.c_void_p()
        assert self.lib.jade_view_create(eng.h, C.byref(self.h)) == 0

    def tick(self):
        n = C.c_int(0)
        assert self.lib.jade_view_tick(self.h, C.byref(n)) == 0, self.lib.jade_last_error(self.eng.h)
        return n.value

    def image(self):
        p, w, h = C.POINTER(C.c_uint32)(), C.c_int(0), C.c_int(0)
        assert self.lib.jade_view_image(self.h, C.byref(p), C.byref(w), C.byref(h)) == 0
        return np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()

    def close(self):
        self.lib.jade_view_destroy(self.h)


def _compare(img, ref, table, label, wrapped_cursor_differs=False):
    assert img.shape == ref.shape, (label, img.shape, ref.shape)
    if wrapped_cursor_differs:
        # The real timerCallback only wraps the cursor column that lands exactly on W (Spectrogram.cpp:715-716); columns
        # W+1.. are written out of bounds there (the stub Image drops them).  The product and the restatement wrap them
        # (DESIGN.md): every red pixel of the reference must be red here, and the extra red ones sit in the first columns.
        assert (img[ref == RED] == RED).all(), f"{label}: red cursor differs"
        extra = (img == RED) & (ref != RED)
        assert not extra[:, 3:].any(), f"{label}: unexpected red pixels"
        ref = np.where(extra, RED, ref).astype(ref.dtype)
    assert np.array_equal(img == RED, ref == RED), f"{label}: red cursor differs"
    diff = img != ref
    # colours come from float32 spectra that differ in the last bits: a few pixels may sit on the other side of a palette
    # edge (tests/parity.py); they must be neighbours in the palette and rare
    assert diff.mean() <= 2e-3, f"{label}: {diff.sum()} of {diff.size} pixels differ"
    if diff.any():
        index = {int(c) | 0xFF000000: i for i, c in reversed(list(enumerate(table)))}
        a = np.array([index[int(c)] for c in img[diff]])
        b = np.array([index[int(c)] for c in ref[diff]])
        assert np.abs(a - b).max() <= 1, f"{label}: palette indices differ by {np.abs(a - b).max()}"


@pytest.mark.parametrize("arm", ["oracle", pytest.param("reference", marks=pytest.mark.skipif(not O.have_ref_spec(), reason="oracle/_ref not built"))])
@pytest.mark.parametrize("N,feed,ch", [(1024, "p50", 2), (2048, "p25", 1), (512, "p100", 2)])
def test_view_follows_reference_timer_callback(N, feed, ch, arm):
    """arm "reference": the comparand is the reference's REAL SpectrogramComponent::timerCallback on its REAL Spectrogram
    (oracle/_ref, Spectrogram.cpp compiled in place), not the restatement."""
    use_ref = arm == "reference"
    fs, mem_s = 48000.0, 0.5
    pct = {"p100": 100, "p50": 50, "p25": 25, "p10": 10}[feed]
    eng = Engine(0, sample_rate=fs, fft_size=N, channels=ch, feed_percent=pct, memory_time_s=mem_s, max_push=N)
    eng.set_palette_scheme("jade", 256)
    eng.set_value_range(-50.0, 50.0)
    spec, pal = O.Spec(use_ref=use_ref), O.Palette(256, O.PAL["jade"])
    spec.set_channels(ch)
    spec.set_samplerate(fs)
    spec.set_memory_time_s(mem_s)
    spec.set_fftsize(N)
    spec.set_feed_percent(O.FEED[feed])
    table = pal.table()
    ov, gv = (O.View(spec, use_ref=True) if use_ref else O.View(spec, pal)), View(eng)
    if use_ref:
        ov.set_scheme(O.PAL["jade"])      # the component's default scheme and range (Spectrogram.cpp:337,342)
        ov.set_color_range(-50.0, 50.0)
    W = spec.memory_size()
    nblocks = 3 * W // spec.feed_blocks() + 5  # wraps the ring several times
    x = signals.streams(1, ch, N * nblocks, fs, kind="mix")[0]
    x *= 20.0
    ticks = 0
    for b in range(nblocks):
        blk = x[:, b * N:(b + 1) * N]
        eng.push(blk)
        spec.process(blk)
        if b == nblocks // 3:          # the "Fix" button: fixed mode with the red cursor
            ov.set_running(False)
            assert eng.lib.jade_view_set_running(gv.h, 0) == 0
        if b == nblocks // 2:          # a range slider moves: full redraw
            ov.set_color_range(-70.0, 20.0)
            assert eng.lib.jade_view_set_value_range(gv.h, -70.0, 20.0) == 0
        if b == (2 * nblocks) // 3:    # back to scrolling
            ov.set_running(True)
            assert eng.lib.jade_view_set_running(gv.h, 1) == 0
        if b % 3 == 1 or b == nblocks - 1:
            no, ng = ov.tick(), gv.tick()
            assert no is None or no == ng, (b, no, ng)  # the real timerCallback returns nothing; the restatement its newVals
            _compare(gv.image(), ov.image(), table, f"block {b}", wrapped_cursor_differs=use_ref)
            ticks += 1
    assert ticks > 6
    gv.close()
    eng.close()


def test_colorbar_matches_the_reference_paint():
    """jade_colorbar against the colourbar the REFERENCE's own paint() drew (Spectrogram.cpp:510-521; tests/golden/paint_ref.npz,
    generated from oracle/_ref): ramp -50..+50 through the current scheme / value range, flipped, | 0xFF000000.  Bit exact."""
    import pathlib

    from jadespectrogram_b200 import engine as E
    gold = np.load(pathlib.Path(__file__).parent / "golden" / "paint_ref.npz")
    for i, case in enumerate(gold["cases"]):
        h, scale, fs, N, scheme, mn, mx = int(case[1]), case[2], case[5], int(case[6]), int(case[7]), case[8], case[9]
        eng = Engine(0, sample_rate=fs, fft_size=N, hop=N // 2, channels=2)
        eng.set_palette_scheme(scheme, 256)
        eng.set_value_range(mn, mx)
        ref = gold[f"c{i}_colorbar"]
        assert E.colorbar_height(h, scale) == len(ref)
        assert np.array_equal(eng.colorbar(len(ref)), ref), f"case {i}"
        eng.close()


def test_view_on_unflipped_and_pooled_engines():
    """The view keeps the reference's orientation (low frequency at the bottom, Spectrogram.cpp:642) whatever row order the
    engine was configured with, and a full redraw works on log max-pool rows (jade_recolor_ring pools the stored dB values)."""
    fs, N, hop = 48000.0, 1024, 512
    x = signals.streams(1, 1, hop * 40, fs, kind="mix")[0]
    imgs = {}
    for flip in (1, 0):
        eng = Engine(0, sample_rate=fs, fft_size=N, hop=hop, channels=1, ring_columns=32, flip_y=flip, row_map="linear_crop",
                     fmin=0.0, fmax=24000.0)
        v = View(eng)
        v.tick()
        for b in range(0, x.shape[1], hop):
            eng.push(x[:, b:b + hop])
            if (b // hop) % 7 == 6:
                v.tick()
        v.tick()
        imgs[flip] = v.image()
        v.close()
        eng.close()
    assert np.array_equal(imgs[0], imgs[1])
    eng = Engine(0, sample_rate=fs, fft_size=N, hop=hop, channels=1, ring_columns=32, row_map="log_maxpool", rows=60, fmin=50.0, fmax=20000.0)
    v = View(eng)
    assert v.tick() > 0            # first tick = full redraw through jade_recolor_ring: used to fail on pooled rows
    for b in range(0, hop * 10, hop):
        eng.push(x[:, b:b + hop])
    v.tick()
    eng.set_value_range(-80.0, 10.0)
    assert eng.lib.jade_view_invalidate(v.h) == 0
    v.tick()
    img = v.image()
    assert img.shape == (60, 32) and (img >> 24 == 0xFF).all()
    # the recoloured ring equals colouring the pooled dB values directly
    pix = np.empty((32, 60), np.uint32)
    assert eng.lib.jade_recolor_ring(eng.h, pix.ctypes.data) == 0
    tot = C.c_int64(0)
    assert eng.lib.jade_ring_info(eng.h, None, None, None, C.byref(tot)) == 0
    total = int(tot.value)  # columns emitted so far (fewer than the ring holds: column c sits in slot c, newest on the right)
    assert 0 < total < 32
    assert np.array_equal(img[:, 32 - total:], pix[:total].T)
    v.close()
    eng.close()
