"""GPU parity tests: the CUDA path, called through the C ABI (libjade_gpu.so), against the CPU oracle."""
import os

import numpy as np
import pytest

import parity
import signals

pytestmark = pytest.mark.gpu

FS = 48000.0


def _oracle_cols(oracle, x, **kw):
    return oracle.render_batch(x, **kw)


CASES = [
    # N, hop, channels, window, mix
    (1024, 512, 1, "hann", "absmean"),        # BASELINE config 1
    (2048, 512, 2, "hann", "absmean"),        # config 2 geometry
    (2048, 256, 1, "hann", "absmean"),        # config 4 geometry
    (16384, 4096, 1, "blackmanharris", "absmean"),  # config 3 geometry
    (64, 16, 1, "rect", "absmean"),
    (128, 32, 2, "hamming", "max"),
    (256, 64, 3, "flattop", "min"),
    (512, 128, 2, "hannpoisson", "left"),
    (512, 100, 2, "hann", "right"),           # odd-ish hop, still even
    (1024, 205, 1, "hann", "absmean"),        # odd hop -> unaligned loads
    (4096, 1024, 2, "hann", "absmean"),
    (8192, 2048, 1, "hann", "max"),
    (32768, 8192, 1, "hann", "absmean"),
    (2048, 512, 1, "hann", "min"),            # Min with one channel clamps at 1e6
    (2048, 510, 2, "hann", "absmean"),        # frames 8- but not 16-byte aligned: the LDG instantiation of the N=2048 kernel
    (2048, 205, 2, "hann", "absmean"),        # odd hop: every column through the guarded instantiation
]


@pytest.mark.parametrize("N,hop,ch,window,mix", CASES)
def test_batch_matches_oracle(gpu_engine_factory, oracle, N, hop, ch, window, mix):
    ncols = 24 if N <= 4096 else 6
    n = hop * ncols
    x = signals.streams(2, ch, n, FS, kind="mix")
    x[1] *= 40.0  # second stream is loud (drives the upper palette clamp and Min's 1e6 start)
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch, window=window, mix_mode=mix)
    assert eng.columns_for(n) == ncols + 1
    pix, db = eng.render_batch(x, want_db=True)
    assert eng.kernel_launches > 0
    B = N // 2 + 1
    for s in range(2):
        odb, opix = oracle.render_batch(x[s], fs=FS, fft_size=N, hop=hop, window=window, mix=mix, ncols=ncols + 1)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")
    assert pix.shape == (2, ncols + 1, B)
    assert (pix >> 24 == 0xFF).all()


def test_n65536_matches_oracle(gpu_engine_factory, oracle):
    N, hop, ncols = 65536, 1024, 5
    n = N + hop * ncols
    x = signals.streams(1, 1, n, 192000.0, kind="mix")
    eng = gpu_engine_factory(sample_rate=192000.0, fft_size=N, hop=hop, channels=1, window="hann")
    pix, db = eng.render_batch(x, first_col=60, ncols=ncols, want_db=True)
    odb, opix = oracle.render_batch(x[0], fs=192000.0, fft_size=N, hop=hop, first_col=60, ncols=ncols)
    parity.check_db(db[0], odb, N)
    parity.check_pixels(pix[0], opix, odb[:, ::-1], -50.0, 50.0, 256)


def test_silence_and_ring_init(gpu_engine_factory):
    eng = gpu_engine_factory(sample_rate=FS, fft_size=1024, hop=512, channels=1, ring_columns=8)
    x = np.zeros((1, 1, 4096), np.float32)
    pix, db = eng.render_batch(x, want_db=True)
    assert np.allclose(db, -110.0, atol=2e-4)  # 10*log10(1e-11)
    ring = eng.read_ring_db()
    assert (ring == -120.0).all()  # Spectrogram.cpp:223


def test_streaming_equals_batch(gpu_engine_factory):
    N, hop, ch = 2048, 512, 2
    n = 512 * 40
    x = signals.streams(1, ch, n, FS, kind="mix")
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch, ring_columns=64, max_push=512)
    bpix, bdb = eng.render_batch(x, want_db=True)
    eng.reset()
    cols_p, cols_d = [], []
    for b in range(40):
        eng.push(x[0][:, b * 512:(b + 1) * 512])
        p, d, first = eng.fetch()
        assert first == len(cols_p) if len(p) else True
        cols_p.extend(p)
        cols_d.extend(d)
    sp, sd = np.array(cols_p), np.array(cols_d)
    assert sp.shape[0] == eng.columns_for(n)
    assert np.array_equal(sd, bdb[0])
    assert np.array_equal(sp, bpix[0])


@pytest.mark.parametrize("N,ch", [(2048, 2), (2048, 1), (1024, 2), (4096, 1)])
def test_streaming_pixels_only_equals_batch(gpu_engine_factory, N, ch):
    """Pixel-only fetches take the polled completion path (armed ring slots, no event wait): same columns as the batch."""
    hop, nblk = 512, 60
    x = signals.streams(1, ch, hop * nblk, FS, kind="mix")
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch, ring_columns=64, max_push=512)
    bpix, _ = eng.render_batch(x)
    eng.reset()
    cols, nxt = [], 0
    for b in range(nblk):
        eng.push(x[0][:, b * hop:(b + 1) * hop])
        p, _, first = eng.fetch(max_cols=8, want_db=False)
        if len(p):
            assert first == nxt
            nxt += len(p)
            cols.extend(p)
    sp = np.array(cols)
    assert sp.shape[0] == eng.columns_for(hop * nblk)
    assert np.array_equal(sp, bpix[0])


def _random_cases(n, seed=20241018):
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n):
        N = int(rng.choice([128, 256, 512, 1024, 2048, 2048, 2048, 4096]))
        hop = int(rng.choice([N // 8, N // 4, N // 2, N // 4 + 2, N // 4 + 1, int(rng.integers(16, N))]))
        ch = int(rng.choice([1, 2, 2, 3, 4]))
        mix = str(rng.choice(["absmean", "absmean", "absmean", "left", "right", "max"]))
        ncols_total = int(rng.integers(12, 40))
        extra = int(rng.integers(0, 7))          # odd lengths: 4-, 8- and unaligned sample counts
        first = int(rng.integers(0, 6))
        cases.append((N, hop, ch, mix, ncols_total, extra, first))
    return cases


@pytest.mark.parametrize("N,hop,ch,mix,ncols_total,extra,first", _random_cases(int(os.environ.get("JADE_RANDOM_CASES", "28"))))
def test_random_geometries_match_oracle(gpu_engine_factory, oracle, N, hop, ch, mix, ncols_total, extra, first):
    """Seeded random geometries: sizes, hops (aligned to 16, 8 bytes, or odd), channel counts / mixes, lengths and
    column windows, so that every routing of launch_stft (staged, LDG, guarded, general) meets the oracle."""
    n = hop * ncols_total + extra
    x = signals.streams(2, ch, n, FS, kind="mix")
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch, window="hann", mix_mode=mix)
    total = eng.columns_for(n)
    ncols = max(1, min(total - first, ncols_total))
    pix, db = eng.render_batch(x, first_col=first, ncols=ncols, want_db=True)
    pix2, _ = eng.render_batch(x, first_col=first, ncols=ncols)
    assert np.array_equal(pix, pix2)  # dB-storing and pixel-only instantiations agree
    for s in range(2):
        odb, opix = oracle.render_batch(x[s], fs=FS, fft_size=N, hop=hop, window="hann", mix=mix, first_col=first, ncols=ncols)
        parity.check_db(db[s], odb, N, f"stream {s}")
        parity.check_pixels(pix[s], opix, odb[:, ::-1], -50.0, 50.0, 256, f"stream {s}")


@pytest.mark.parametrize("seed,N,hop,ch", [(1, 2048, 512, 2), (2, 2048, 256, 1), (3, 1024, 512, 2), (4, 512, 128, 1),
                                            (5, 4096, 1024, 2), (6, 2048, 510, 2), (7, 2048, 333, 1)])
def test_streaming_random_pushes_equal_batch(gpu_engine_factory, seed, N, hop, ch):
    """Pushes of random sizes (1 .. max_push samples: history slides, staging-slot reuse, columns appearing several at a
    time), fetches alternating between the polled pixel-only path and the dB path: bit-identical to the batch render."""
    rng = np.random.default_rng(seed)
    n = hop * 90 + int(rng.integers(0, hop))
    x = signals.streams(1, ch, n, FS, kind="mix")
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch, ring_columns=128, max_push=1500)
    bpix, bdb = eng.render_batch(x, want_db=True)
    eng.reset()
    pos, nxt, k = 0, 0, 0
    while pos < n:
        m = int(min(n - pos, rng.integers(1, 1501)))
        eng.push(x[0][:, pos:pos + m])
        pos += m
        want_db = (k % 3 == 2)
        k += 1
        p, d, first = eng.fetch(max_cols=64, want_db=want_db)
        if len(p):
            assert first == nxt
            assert np.array_equal(p, bpix[0][nxt:nxt + len(p)])
            if want_db:
                assert np.array_equal(d, bdb[0][nxt:nxt + len(p)])
            nxt += len(p)
    assert nxt == eng.columns_for(n)


def test_block_emit_mode_matches_reference_counts(gpu_engine_factory, oracle):
    """Reference emission pattern: feed 50 % -> 2 columns per N-sample block, newest column ends at (b+1)N - hop."""
    N = 1024
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, feed_percent=50, channels=1, memory_time_s=1.0)
    sp = oracle.Spec()
    sp.set_channels(1)
    sp.set_samplerate(FS)
    sp.set_memory_time_s(1.0)
    sp.set_fftsize(N)
    sp.set_feed_percent(oracle.FEED["p50"])
    assert eng.W == sp.memory_size()
    x = signals.streams(1, 1, N * 12, FS, kind="mix")[0]
    mem = np.zeros((sp.memory_size(), sp.spectrum_size()), np.float32)
    total = 0
    for b in range(12):
        blk = x[:, b * N:(b + 1) * N]
        sp.process(blk)
        eng.push(blk)
        p, d, first = eng.fetch()
        assert len(p) == 2 and first == total
        newv, pos = sp.get_mem(mem)
        if b > 0:
            assert newv == 2
        parity.check_db(d, mem[[(pos - 2) % eng.W, (pos - 1) % eng.W]], N, f"block {b}")
        total += 2


@pytest.mark.skipif(not __import__("oracle_lib").have_ref_spec(), reason="oracle/_ref (compiled reference) not built")
@pytest.mark.parametrize("N,feed,ch,window", [(2048, 25, 2, "hann"), (1024, 50, 1, "hann"), (2048, 10, 2, "blackmanharris"),
                                              (512, 100, 2, "hamming"), (4096, 25, 1, "flattop"), (2048, 50, 2, "hannpoisson")])
def test_batch_matches_the_real_reference_class(gpu_engine_factory, oracle, N, feed, ch, window):
    """jade_render_batch against the reference's REAL Spectrogram + CColorPalette (Spectrogram.cpp / CColorpalette.cpp compiled
    in place into oracle/_ref/libjade_ref.so; only the FFT inside is the stand-in): the reference class is fed N-sample blocks
    (processSynchronBlock, all four feed percentages incl. the non-uniform 10 % hop), its ring is read with getMem and coloured
    with its own getRGBColor; the GPU renders the same samples in one batch call with the reference block geometry."""
    fs, nblocks = FS, 14
    x = signals.streams(1, ch, N * nblocks, fs, kind="mix")[0]
    x *= 8.0
    ref = oracle.Spec(use_ref=True)
    ref.set_channels(ch)
    ref.set_samplerate(fs)
    ref.set_memory_time_s(30.0)  # a ring longer than the run: slot = column index
    ref.set_fftsize(N)
    ref.set_feed_percent(oracle.FEED[f"p{feed}"])
    ref.set_window(oracle.WIN[window])
    W, B = ref.memory_size(), ref.spectrum_size()
    mem = np.zeros((W, B), np.float32)
    ref.get_mem(mem)
    for b in range(nblocks):
        ref.process(x[:, b * N:(b + 1) * N])
    newv, pos = ref.get_mem(mem)
    ncols = nblocks * ref.feed_blocks()
    assert newv == ncols == pos
    rdb = mem[:ncols]
    pal = oracle.Palette(256, oracle.PAL["jade"], use_ref=True)
    pal.set_value_range(-50.0, 50.0)
    rpix = (pal.lookup(rdb[:, ::-1].reshape(-1)).reshape(ncols, B).astype(np.uint32)) | np.uint32(0xFF000000)

    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, feed_percent=feed, channels=ch, window=window)
    assert eng.columns_for(N * nblocks) == ncols
    pix, db = eng.render_batch(x[None], want_db=True)
    parity.check_db(db[0], rdb, N, "vs reference class")
    parity.check_pixels(pix[0], rpix, rdb[:, ::-1], -50.0, 50.0, 256, "vs reference class")
