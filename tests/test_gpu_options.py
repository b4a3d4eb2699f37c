"""GPU tests of the engine options beyond the headline path: row maps, pixel formats, value range / palette changes and
the re-colour kernel, precise dB, power_scale, ring overflow, multi-engine batch, device-pointer API, properties at
full BASELINE sizes."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import parity
import signals

pytestmark = pytest.mark.gpu
FS = 48000.0


def _ref_pixels(db_rows, scheme, n, mn, mx, invert=False):
    p = O.Palette(n, O.PAL[scheme])
    if invert:
        p.set_invert(1)
        p.set_color_scheme(O.PAL[scheme])
    p.set_value_range(mn, mx)
    return p.lookup(db_rows).astype(np.uint32) | np.uint32(0xFF000000)


def test_row_maps_and_orientation(gpu_engine_factory):
    N, hop = 2048, 512
    x = signals.streams(1, 1, hop * 20, FS)
    odb, _ = O.render_batch(x[0], fft_size=N, hop=hop, ncols=21)
    B = N // 2 + 1
    # un-flipped identity rows
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, flip_y=0)
    pix, db = eng.render_batch(x, want_db=True)
    parity.check_db(db[0], odb, N)
    parity.check_pixels(pix[0], _ref_pixels(odb, "jade", 256, -50, 50), odb, -50, 50, 256)
    # linear crop (the reference's paint() maths), flipped
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, row_map="linear_crop", fmin=200.0, fmax=12000.0)
    lo, hi = C.c_int(), C.c_int()
    eng.lib.jade_linear_crop(FS, B, 200.0, 12000.0, C.byref(lo), C.byref(hi))
    assert eng.R == hi.value - lo.value
    pix, _ = eng.render_batch(x)
    ref = _ref_pixels(odb, "jade", 256, -50, 50)[:, lo.value:hi.value][:, ::-1]
    parity.check_pixels(pix[0], ref, odb[:, lo.value:hi.value][:, ::-1], -50, 50, 256)
    # log max-pool
    R = 200
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, row_map="log_maxpool", rows=R, fmin=30.0, fmax=20000.0)
    blo, bhi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    eng.lib.jade_log_rows(FS, N, R, 30.0, 20000.0, blo.ctypes.data, bhi.ctypes.data)
    pix, _ = eng.render_batch(x)
    pooled = np.stack([odb[:, blo[r]:bhi[r]].max(axis=1) for r in range(R)], axis=1)
    parity.check_pixels(pix[0], _ref_pixels(pooled, "jade", 256, -50, 50)[:, ::-1], pooled[:, ::-1], -50, 50, 256)


def test_cfg5_geometry_log_rows_n65536(gpu_engine_factory):
    """BASELINE config 5 geometry (mono 192 kHz, FFT 65536, hop 1024, 1080 log rows) on a short excerpt."""
    fs, N, hop, R = 192000.0, 65536, 1024, 1080
    x = signals.streams(1, 1, N + hop * 8, fs)
    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, hop=hop, channels=1, row_map="log_maxpool", rows=R, fmin=20.0, fmax=96000.0)
    assert eng.kernel_name == "pkcl3<65536>"  # one contributing channel: three register passes per CTA (jade_pk_cluster3.cuh)
    pix, db = eng.render_batch(x, first_col=64, ncols=6, want_db=True)
    odb, _ = O.render_batch(x[0], fs=fs, fft_size=N, hop=hop, first_col=64, ncols=6)
    parity.check_db(db[0], odb, N)
    blo, bhi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    eng.lib.jade_log_rows(fs, N, R, 20.0, 96000.0, blo.ctypes.data, bhi.ctypes.data)
    pooled = np.stack([odb[:, blo[r]:bhi[r]].max(axis=1) for r in range(R)], axis=1)
    parity.check_pixels(pix[0], _ref_pixels(pooled, "jade", 256, -50, 50)[:, ::-1], pooled[:, ::-1], -50, 50, 256)


@pytest.mark.parametrize("scheme,n,invert,rng", [("viridis", 256, False, (-80.0, 0.0)), ("rainbow", 64, True, (-100.0, -10.0)),
                                                 ("hot", 1024, False, (20.0, -60.0)), ("bw", 2, False, (-30.0, -30.0 + 1e-3)),
                                                 ("mono", 7, False, (-50.0, 50.0)), ("plasma", 255, True, (-120.0, 60.0)),
                                                 # 256 colours, `>= m_Max` entry != entry 255 (index of 0.9999 * 59 is 254): not a u8 palette
                                                 ("rainbow", 256, False, (58.0, 59.0))])
def test_palettes_and_ranges(gpu_engine_factory, scheme, n, invert, rng):
    N, hop = 1024, 256
    x = signals.streams(1, 2, hop * 16, FS)
    x *= 30.0
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=2)
    # the drop-in class's way: constructor table, then invert + rebuild in place
    tab = np.zeros(n, np.int32)
    eng.lib.jade_palette_build(O.PAL[scheme], n, 0, tab.ctypes.data)
    if invert:
        eng.lib.jade_palette_build(O.PAL[scheme], n, 1, tab.ctypes.data)
    eng.set_palette(tab)
    eng.set_value_range(*rng)
    pix, db = eng.render_batch(x, want_db=True)
    odb, _ = O.render_batch(x[0], fft_size=N, hop=hop, ncols=17)
    parity.check_db(db[0], odb, N)
    ref = _ref_pixels(odb, scheme, n, rng[0], rng[1], invert)[:, ::-1]
    parity.check_pixels(pix[0], ref, odb[:, ::-1], min(rng), max(rng), n)
    assert eng.lookup_color(odb[3, 40]) == int(ref[3, N // 2 - 40] & 0xFFFFFF)
    pix2, _ = eng.render_batch(x)  # the pixel-only instantiation
    assert np.array_equal(pix2, pix)


@pytest.mark.parametrize("N,hop,ch", [(2048, 512, 2), (16384, 4096, 1), (2048, 256, 1), (1024, 512, 1)])
def test_u8_palette_kernels_match_the_clamping_ones(gpu_engine_factory, N, hop, ch):
    """256-colour palettes whose `>= m_Max` colour is the last one take kernels in which the saturating float -> u8 conversion
    is the whole index clamp (KParams::pal_u8; separate pixel-only instantiations for N = 2048 stereo and N = 16384): same
    pixels as the instantiations with the integer clamp (the dB-storing ones) and, for another range, as the oracle."""
    x = signals.streams(2, ch, hop * 40 + N, FS)
    x *= 30.0
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=ch)
    for rng in ((-50.0, 50.0), (0.0, 60.0), (58.0, 59.0)):
        eng.set_value_range(*rng)
        pix_db, db = eng.render_batch(x, want_db=True)
        pix, _ = eng.render_batch(x)
        assert np.array_equal(pix, pix_db), rng
    odb, _ = O.render_batch(x[1], fft_size=N, hop=hop, ncols=pix.shape[1])
    parity.check_pixels(pix[1], _ref_pixels(odb, "jade", 256, 58.0, 59.0)[:, ::-1], odb[:, ::-1], 58.0, 59.0, 256)


def test_rgba8_byte_order(gpu_engine_factory):
    x = signals.streams(1, 1, 512 * 8, FS)
    a = gpu_engine_factory(sample_rate=FS, fft_size=1024, hop=512, channels=1, pixel_format="argb32")
    pa, _ = a.render_batch(x)
    b = gpu_engine_factory(sample_rate=FS, fft_size=1024, hop=512, channels=1, pixel_format="rgba8")
    pb, _ = b.render_batch(x)
    by = pb.view(np.uint8).reshape(pb.shape + (4,))
    assert np.array_equal(by[..., 0], (pa >> 16) & 255) and np.array_equal(by[..., 1], (pa >> 8) & 255)
    assert np.array_equal(by[..., 2], pa & 255) and (by[..., 3] == 255).all()


def test_precise_db_and_power_scale(gpu_engine_factory):
    N, hop = 512, 128
    x = signals.streams(1, 1, hop * 12, FS)
    odb, _ = O.render_batch(x[0], fft_size=N, hop=hop, ncols=13)
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, db_precise=1)
    _, db = eng.render_batch(x, want_db=True)
    parity.check_db(db[0], odb, N)
    assert (db[0][0] == np.float32(10.0 * np.log10(np.float64(np.float32(1e-11))))).all()  # silence is bit-exact
    # power_scale = 1/N^2 (a "normalised FFT" convention) shifts every bin by -20*log10(N) dB
    eng2 = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, power_scale=1.0 / (N * N))
    _, db2 = eng2.render_batch(x, want_db=True)
    loud = odb > -10  # after the -54 dB shift these are still far above the 1e-11 (-110 dB) floor
    assert loud.sum() > 100
    assert np.allclose(db2[0][loud], odb[loud] - 20 * np.log10(N), atol=2e-3)


def test_ring_overflow_recolor_and_fetch_limits(gpu_engine_factory):
    N, hop, W = 256, 64, 16
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1, ring_columns=W, max_push=4096)
    x = signals.streams(1, 1, 4096 * 3, FS)[0]
    full, _ = eng.render_batch(x[None], want_db=False)
    eng.reset()
    eng.push(x[:, :4096])  # 65 columns at once: only the newest 16 survive
    pix, db, first = eng.fetch()
    ncol = eng.columns_for(4096)
    assert len(pix) == W and first == ncol - W
    assert np.array_equal(pix, full[0][ncol - W:ncol])
    eng.push(x[:, 4096:4096 + 640])
    pix2, db2, first2 = eng.fetch(max_cols=4)
    assert len(pix2) == 4 and first2 == eng.columns_for(4096 + 640) - 4
    # re-colour the ring with another palette / range == render with that palette
    eng.set_palette_scheme("viridis", 256)
    eng.set_value_range(-90.0, 10.0)
    rec = eng.recolor_ring()
    ringdb = eng.read_ring_db()
    ref = _ref_pixels(ringdb, "viridis", 256, -90.0, 10.0)[:, ::-1]
    assert np.array_equal(rec, ref)
    assert eng.ring_info()[3] == eng.columns_for(4096 + 640)


def test_multi_engine_batch_equals_single(gpu_engine_factory):
    """jade_render_batch_multi range-partitions streams / columns with no exchange; two engines on one GPU must give the
    bit-identical image (the multi-GPU path with the halo re-read)."""
    from jadespectrogram_b200 import Engine
    N, hop = 1024, 256
    for nstreams in (5, 1):
        x = signals.streams(nstreams, 2, hop * 40, FS)
        e1 = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=2)
        e2 = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=2)
        e3 = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=2)
        single, _ = e1.render_batch(x)
        ncols = single.shape[1]
        out = np.zeros_like(single)
        hs = (C.c_void_p * 2)(e2.h, e3.h)
        rc = e1.lib.jade_render_batch_multi(hs, 2, x.ctypes.data, nstreams, x.shape[2], 0, ncols, out.ctypes.data, None)
        assert rc == 0
        assert np.array_equal(out, single)


def test_device_pointer_api_matches_host_api(gpu_engine_factory):
    import torch
    N, hop, S = 2048, 512, 6
    x = signals.streams(S, 2, hop * 64, FS)
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=2)
    host, _ = eng.render_batch(x)
    ncols = host.shape[1]
    d_in = torch.from_numpy(x).cuda()
    d_pix = torch.zeros((S, ncols, N // 2 + 1), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng.render_device(d_in.data_ptr(), S, x.shape[2], 2 * x.shape[2], x.shape[2], 0, ncols, d_pix.data_ptr(), None, st.cuda_stream)
    st.synchronize()
    assert np.array_equal(d_pix.cpu().numpy().view(np.uint32), host)
    assert eng.last_kernel_seconds() > 0


def test_full_size_properties_cfg4(gpu_engine_factory):
    """BASELINE config 4 geometry at scale (many mono streams, FFT 2048, hop 256): size-independent properties --
    linearity in dB (x -> 10x adds 20 dB), stream independence, determinism."""
    import torch
    N, hop, S, n = 2048, 256, 256, 48000 * 4
    eng = gpu_engine_factory(sample_rate=FS, fft_size=N, hop=hop, channels=1)
    ncols = eng.columns_for(n)
    d_in = torch.empty((S, 1, n), dtype=torch.float32, device="cuda")
    eng.synth_device(d_in.data_ptr(), S, 1, n, n, n, kind="mix")
    d_db = torch.empty((S, ncols, N // 2 + 1), dtype=torch.float32, device="cuda")
    d_pix = torch.empty((S, ncols, N // 2 + 1), dtype=torch.int32, device="cuda")
    eng.render_device(d_in.data_ptr(), S, n, n, n, 0, ncols, d_pix.data_ptr(), d_db.data_ptr())
    eng.sync()
    a = d_db.clone()
    p1 = d_pix.clone()
    eng.render_device(d_in.data_ptr(), S, n, n, n, 0, ncols, d_pix.data_ptr(), d_db.data_ptr())
    eng.sync()
    assert torch.equal(a, d_db) and torch.equal(p1, d_pix)  # deterministic
    d_in *= 10.0
    eng.render_device(d_in.data_ptr(), S, n, n, n, 0, ncols, d_pix.data_ptr(), d_db.data_ptr())
    eng.sync()
    # linearity with the float32-FFT noise-floor term of the parity criterion (bins ~60 dB under the frame peak move)
    # (column 0 is the all-zero pre-roll: its -110 dB floor does not scale)
    parity.check_db(d_db[:8, 1:].cpu().numpy(), (a[:8, 1:] + 20.0).cpu().numpy(), N, "x10 -> +20 dB")
    strong = (a > (a.amax(dim=-1, keepdim=True) - 30.0)) & (a > -60.0)
    assert torch.allclose(d_db[strong], a[strong] + 20.0, atol=2e-3)
    # a stream rendered alone equals its slice of the batch
    one = torch.empty((1, ncols, N // 2 + 1), dtype=torch.float32, device="cuda")
    eng.render_device(d_in[17].data_ptr(), 1, n, n, n, 0, ncols, d_pix.data_ptr(), one.data_ptr())
    eng.sync()
    assert torch.equal(one[0], d_db[17])
    # first column is the zero pre-roll; the synthetic signal check against the oracle on one stream prefix
    assert torch.allclose(d_db[:, 0], torch.full_like(d_db[:, 0], -110.0), atol=2e-4)
    xs = d_in[3, :, :hop * 8].cpu().numpy()
    odb, _ = O.render_batch(xs, fft_size=N, hop=hop, ncols=9)
    parity.check_db(d_db[3, :9].cpu().numpy(), odb, N)
