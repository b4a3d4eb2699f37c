"""Kernel logic on the CPU SIMT emulator (tests/emu): the product's kernel source (jade_kernels.cuh) compiled with
-DJADE_EMU and executed with one OS thread per CUDA thread, against the oracle.  This is a debugging aid for the
GPU-less build container -- it proves index arithmetic, shared-memory layouts and barriers, not performance -- and it is
not a product path (the product library contains no host implementation of these kernels)."""
import numpy as np
import pytest

import emu_lib as E
import oracle_lib as O
import parity
import signals
from jadespectrogram_b200._capi import MIX, ROWS, WIN, JadeConfig

pytestmark = pytest.mark.timeout(300)


def _cfg(N, hop, ch, window, mix, **kw):
    c = JadeConfig()
    c.sample_rate = 48000.0
    c.fft_size, c.hop, c.frames_per_block, c.block_stride, c.preroll = N, hop, 1, hop, -1
    c.window, c.channels, c.mix_mode = WIN[window], ch, MIX[mix]
    c.row_map, c.flip_y, c.power_scale = 0, 1, 1.0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


CASES = [(64, 16, 1, "rect", "absmean"), (128, 32, 2, "hann", "absmean"), (256, 64, 1, "hamming", "absmean"),
         (512, 128, 3, "flattop", "absmean"), (1024, 512, 1, "hann", "absmean"), (2048, 512, 2, "hann", "absmean"),
         (2048, 256, 1, "hann", "min"), (1024, 205, 2, "hannpoisson", "max"), (2048, 512, 2, "hann", "right"),
         (512, 128, 2, "hann", "absmean"), (512, 100, 1, "blackmanharris", "left"), (256, 63, 1, "hann", "absmean"),
         (1024, 256, 4, "hann", "absmean"), (4096, 1024, 1, "hann", "absmean"), (16384, 4096, 1, "blackmanharris", "absmean"), (65536, 8192, 1, "hann", "absmean")]


@pytest.mark.parametrize("N,hop,ch,window,mix", CASES)
def test_emulated_kernels_match_oracle(N, hop, ch, window, mix):
    ncols = 9 if N <= 2048 else 3
    x = signals.streams(2 if N <= 2048 else 1, ch, hop * (ncols - 1) + 64, 48000.0)
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, ch, window, mix), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=2)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window=window, mix=mix, ncols=ncols)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")


@pytest.mark.parametrize("ch,mix", [(2, "absmean"), (1, "absmean"), (2, "right"), (4, "absmean")])
@pytest.mark.parametrize("variant", ["pair2", "single"])
def test_emulated_pk2048_pair_and_single(monkeypatch, ch, mix, variant):
    """N = 2048, TMA-staged interior frames: the two-real-transforms stereo kernel (JADE_EMU_PAIR2) and the one-transform
    kernel (JADE_EMU_NOPAIR, and always for one or four contributing channels) against the oracle, and bit-identical to
    each other."""
    monkeypatch.delenv("JADE_EMU_NOPAIR", raising=False)
    monkeypatch.delenv("JADE_EMU_PAIR2", raising=False)
    monkeypatch.setenv("JADE_EMU_NOPAIR" if variant == "single" else "JADE_EMU_PAIR2", "1")
    N, hop, ncols = 2048, 512, 29
    x = signals.streams(2, ch, hop * (ncols - 1) + 64, 48000.0)
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, ch, "hann", mix), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window="hann", mix=mix, ncols=ncols)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")
    key = (ch, mix)
    prev = _PAIR_RESULTS.setdefault(key, (db, pix))
    assert np.array_equal(prev[0], db) and np.array_equal(prev[1], pix)


@pytest.mark.parametrize("window", ["hann", "blackmanharris", "hannpoisson"])
@pytest.mark.parametrize("want_db", [True, False])
def test_emulated_pkz2048_complex_stereo(monkeypatch, window, want_db):
    """N = 2048, AbsMean over two channels -- the product route: both channels as ONE 2048-point complex transform
    (stft_pkz2048_kernel, |X_L|^2 + |X_R|^2 = (|Z[k]|^2 + |Z[N-k]|^2) / 2).  Against the oracle over a chain of frames per
    warp (TMA staging of the next frame, rotated loop), and the TMA-staged instantiation bit-identical to the guarded one
    (streaming == batch == sharded relies on it)."""
    for k in ("JADE_EMU_NOPAIR", "JADE_EMU_PAIR2", "JADE_EMU_FORCE_GUARD", "JADE_EMU_RING"):
        monkeypatch.delenv(k, raising=False)
    N, hop, ncols = 2048, 512, 53
    x = signals.streams(2, 2, hop * (ncols - 1) + 64, 48000.0, kind="mix")
    x[1, 1] *= 1e-3  # a nearly silent right channel next to a loud left one
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, 2, window, "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1, want_db=want_db)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window=window, mix="absmean", ncols=ncols)
        if want_db:
            parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")
    monkeypatch.setenv("JADE_EMU_FORCE_GUARD", "1")
    gdb, gpix = E.render(_cfg(N, hop, 2, window, "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=2, want_db=True)
    assert np.array_equal(gpix, pix)
    if want_db:
        assert np.array_equal(gdb, db)
    # the long-run instantiation (contiguous columns per warp, samples in a tensor-memory ring of four chunks): runs that start
    # in the middle of a stream and cross into the next one
    monkeypatch.delenv("JADE_EMU_FORCE_GUARD")
    monkeypatch.setenv("JADE_EMU_RING", "1")
    rdb, rpix = E.render(_cfg(N, hop, 2, window, "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1, want_db=want_db)
    assert np.array_equal(rpix, pix)
    if want_db:
        assert np.array_equal(rdb, db)


@pytest.mark.parametrize("want_db", [True, False])
def test_emulated_pk3_16384(monkeypatch, want_db):
    """N = 16384, one contributing channel -- the product route: three register passes 32 x 16 x 16 by one CTA, tables in
    tensor memory, split from registers (stft_pk3_kernel).  Interior frames through the TMA-staged instantiation (a chain of
    frames per CTA: the next frame is staged while the current one is in pass 3), boundary frames through the guarded one;
    against the oracle, and staged == guarded bit for bit (streaming == batch == sharded relies on it)."""
    for k in ("JADE_EMU_FORCE_GUARD", "JADE_EMU_PKCTA"):
        monkeypatch.delenv(k, raising=False)
    N, hop, ncols = 16384, 4096, 9
    x = signals.streams(2, 2, hop * (ncols - 1) - 4096, 96000.0, kind="mix")  # the last columns run past the end
    pal = O.Palette(256, O.PAL["jade"]).table()
    cfg = _cfg(N, hop, 2, "blackmanharris", "right")
    db, pix = E.render(cfg, pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=2, want_db=want_db)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window="blackmanharris", mix="right", ncols=ncols)
        if want_db:
            parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")
    monkeypatch.setenv("JADE_EMU_FORCE_GUARD", "1")
    gdb, gpix = E.render(cfg, pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=3, want_db=True)
    assert np.array_equal(gpix, pix)
    if want_db:
        assert np.array_equal(gdb, db)


@pytest.mark.parametrize("hop", [256, 512])
def test_emulated_pk2048_mono_ring(monkeypatch, hop):
    """N = 2048, one contributing channel, hop 256 / 512: the long-run instantiations of stft_pk2048_kernel (PK_LD_RING4 / RING8:
    contiguous columns per warp, the frame's values in a tensor-memory ring of 8 / 4 chunks, only the last chunk staged for a
    frame that continues its predecessor) are bit-identical to the round-robin instantiation, over runs that start in the
    middle of a stream and cross into the next one."""
    for k in ("JADE_EMU_NOPAIR", "JADE_EMU_PAIR2", "JADE_EMU_FORCE_GUARD", "JADE_EMU_RING"):
        monkeypatch.delenv(k, raising=False)
    N, ncols = 2048, 61
    x = signals.streams(2, 2, hop * (ncols - 1) + 64, 48000.0, kind="mix")
    pal = O.Palette(256, O.PAL["jade"]).table()
    cfg = _cfg(N, hop, 2, "hamming", "left")
    db, pix = E.render(cfg, pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window="hamming", mix="left", ncols=ncols)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")
    monkeypatch.setenv("JADE_EMU_RING", "1")
    rdb, rpix = E.render(cfg, pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1)
    assert np.array_equal(rpix, pix) and np.array_equal(rdb, db)


_PAIR_RESULTS = {}


@pytest.mark.parametrize("N,hop,ch,ncols", [(1024, 512, 1, 70), (1024, 256, 2, 45), (256, 64, 2, 150), (128, 32, 1, 301)])
def test_emulated_pksmall_staging_chain(N, hop, ch, ncols):
    """N <= 1024 packed kernels with several column groups per warp: the TMA staging of the next channel / next group
    (F frames per warp on one mbarrier), including a padded last group (odd column counts)."""
    x = signals.streams(2, ch, hop * (ncols - 1) + 64, 48000.0)
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, ch, "hann", "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window="hann", mix="absmean", ncols=ncols)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")


@pytest.mark.parametrize("extra,want_db", [(0, True), (0, False), (2, True), (1, True)])
def test_emulated_pk2048_prefetch_chain(extra, want_db):
    """N = 2048 with several frames per warp: the cp.async staging of the next channel / next frame (16-byte aligned
    input, extra = 0), the LDG-to-register instantiation (8-byte aligned only, extra = 2) and the guarded one (odd
    length, extra = 1) must give the same columns as the oracle, for both the pixel-only and the dB-storing forms."""
    N, hop, ch, ncols = 2048, 512, 2, 30
    x = signals.streams(2, ch, hop * (ncols - 1) + 64 + extra, 48000.0)
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, ch, "hann", "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1, want_db=want_db)
    for s in range(x.shape[0]):
        odb, opix = O.render_batch(x[s], fft_size=N, hop=hop, window="hann", mix="absmean", ncols=ncols)
        if want_db:
            parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")


@pytest.mark.parametrize("N,hop,ch,extra", [(4096, 1024, 2, 0), (8192, 2048, 1, 0), (4096, 1024, 1, 2)])
def test_emulated_pkcta_frame_staging(N, hop, ch, extra):
    """N >= 4096 packed kernels, several frames per CTA: interior frames are staged into the row matrix by the TMA engine
    (next channel / next frame while the epilogue runs), boundary frames and 8-byte-aligned inputs (extra = 2) take the
    global-load paths; the sequence mixes both."""
    ncols = 11
    x = signals.streams(1, ch, hop * (ncols - 1) + 64 + extra, 48000.0)
    pal = O.Palette(256, O.PAL["jade"]).table()
    db, pix = E.render(_cfg(N, hop, ch, "hann", "absmean"), pal, -50.0, 50.0, x, 0, ncols, N // 2 + 1, grid=1)
    odb, opix = O.render_batch(x[0], fft_size=N, hop=hop, window="hann", mix="absmean", ncols=ncols)
    parity.check_db(db[0], odb, N)
    parity.check_pixels(pix[0], opix, odb[:, ::-1], -50.0, 50.0, 256)


def test_emulated_general_epilogue_options():
    N, hop = 1024, 256
    x = signals.streams(1, 2, hop * 8, 48000.0)
    pal = O.Palette(64, O.PAL["viridis"]).table()
    odb, _ = O.render_batch(x[0], fft_size=N, hop=hop, ncols=8)
    # precise dB reproduces the oracle's double log10 on identical power up to float FFT differences; un-flipped rows
    db, pix = E.render(_cfg(N, hop, 2, "hann", "absmean", db_precise=1, flip_y=0), pal, -80.0, 0.0, x, 0, 8, N // 2 + 1)
    parity.check_db(db[0], odb, N)
    ref = O.Palette(64, O.PAL["viridis"])
    ref.set_value_range(-80.0, 0.0)
    parity.check_pixels(pix[0], ref.lookup(odb).astype(np.uint32) | np.uint32(0xFF000000), odb, -80.0, 0.0, 64)
    # linear crop: rows are the bins the reference's paint() would show
    c = _cfg(N, hop, 2, "hann", "absmean", row_map=ROWS["linear_crop"], fmin=1000.0, fmax=8000.0)
    lo, hi = 43, 171  # int(2*8000/48000*513+.5)=171 ; interval=int(171.0-21.375+.5)=150 -> lo = 21? computed below
    from jadespectrogram_b200 import _capi
    import ctypes as C
    a, b = C.c_int(), C.c_int()
    _capi.load().jade_linear_crop(48000.0, N // 2 + 1, 1000.0, 8000.0, C.byref(a), C.byref(b))
    lo, hi = a.value, b.value
    db, pix = E.render(c, pal, -80.0, 0.0, x, 0, 8, hi - lo)
    full = ref.lookup(odb).astype(np.uint32) | np.uint32(0xFF000000)
    parity.check_pixels(pix[0], full[:, lo:hi][:, ::-1], odb[:, lo:hi][:, ::-1], -80.0, 0.0, 64)
    # log max-pool rows: dB of the band maximum
    R = 40
    c = _cfg(N, hop, 2, "hann", "absmean", row_map=ROWS["log_maxpool"], rows=R, fmin=50.0, fmax=20000.0)
    blo, bhi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    _capi.load().jade_log_rows(48000.0, N, R, 50.0, 20000.0, blo.ctypes.data, bhi.ctypes.data)
    db, pix = E.render(c, pal, -80.0, 0.0, x, 0, 8, R)
    pooled = np.stack([odb[:, blo[r]:bhi[r]].max(axis=1) for r in range(R)], axis=1)
    parity.check_pixels(pix[0], (ref.lookup(pooled).astype(np.uint32) | np.uint32(0xFF000000))[:, ::-1], pooled[:, ::-1],
                        -80.0, 0.0, 64)


def test_emulated_n65536_log_rows_without_db():
    """BASELINE config 5 path: pixels only, log max-pool rows -> the warp-per-row pooled epilogue (shuffle max)."""
    from jadespectrogram_b200 import _capi
    N, hop, R, fs = 65536, 1024, 96, 192000.0
    x = signals.streams(1, 1, N + hop * 3, fs)
    pal = O.Palette(256, O.PAL["jade"]).table()
    c = _cfg(N, hop, 1, "hann", "absmean", row_map=ROWS["log_maxpool"], rows=R, fmin=20.0, fmax=96000.0)
    c.sample_rate = fs
    _, pix = E.render(c, pal, -50.0, 50.0, x, 64, 3, R, grid=1, want_db=False)
    odb, _ = O.render_batch(x[0], fs=fs, fft_size=N, hop=hop, first_col=64, ncols=3)
    blo, bhi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    _capi.load().jade_log_rows(fs, N, R, 20.0, 96000.0, blo.ctypes.data, bhi.ctypes.data)
    pooled = np.stack([odb[:, blo[r]:bhi[r]].max(axis=1) for r in range(R)], axis=1)
    ref = O.Palette(256, O.PAL["jade"])
    ref.set_value_range(-50.0, 50.0)
    parity.check_pixels(pix[0], (ref.lookup(pooled).astype(np.uint32) | np.uint32(0xFF000000))[:, ::-1], pooled[:, ::-1],
                        -50.0, 50.0, 256)
