"""BASELINE.json configs 3, 4 and 5 at their FULL sizes on one B200, through size-independent properties.

The oracle cannot run hours of audio in seconds, so each full-size run is checked by
  * spot parity  : a few (stream, column-window) excerpts are copied to the host and compared with the CPU oracle under
                   the north-star tolerances (tests/parity.py);
  * shard invariance : column ranges rendered separately (the multi-GPU partition, incl. its N-hop input halo) are
                   bit-identical to the same columns of the whole rendering -- a checksum of checksums over all slabs;
  * pre-roll     : column 0 of every stream analyses only the reference's N zeros (Spectrogram.cpp:233) => -110 dB colour;
  * determinism  : a second pass reproduces the checksums.
Inputs are generated on the device (jade_synth_device); nothing here reads /root/reference.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import parity

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _floor_pixel(eng):
    return np.uint32(eng.lookup_color(-110.0)) | np.uint32(0xFF000000)


def _checksum(t):
    import torch
    # order-sensitive 64-bit checksum computed on the device
    flat = t.reshape(-1)
    total, step = 0, 1 << 27
    for i in range(0, flat.numel(), step):
        v = flat[i:i + step].to(torch.int64) & 0xFFFFFFFF
        w = (torch.arange(i, i + v.numel(), device=v.device, dtype=torch.int64) % 65521) + 1
        total = (total + int((v * w).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return total


def _spot_check(eng, d_stream, n, N, hop, c0, k, pix_dev_rows, rows_from_db, label):
    """d_stream: device tensor [channels][n]; pix_dev_rows: device pixels [k][R] of columns c0..c0+k-1."""
    a = max(0, c0 * hop - N)
    b = min(n, (c0 + k - 1) * hop)
    xs = d_stream[:, a:b].cpu().numpy()
    first = c0 - a // hop
    odb, _ = O.render_batch(xs, fs=eng.cfg.sample_rate, fft_size=N, hop=hop, window=_win_name(eng), first_col=first, ncols=k)
    ref_rows_db = rows_from_db(odb)
    p = O.Palette(256, O.PAL["jade"])
    p.set_value_range(-50.0, 50.0)
    ref_pix = p.lookup(ref_rows_db).astype(np.uint32) | np.uint32(0xFF000000)
    got = pix_dev_rows.cpu().numpy().view(np.uint32)
    return parity.check_pixels(got, ref_pix, ref_rows_db, -50.0, 50.0, 256, label)


def _win_name(eng):
    return [k for k, v in O.WIN.items() if v == eng.cfg.window][0]


@pytest.mark.parametrize("ch,hop,kernel", [(2, 512, "pkz2048"), (1, 512, "pk2048"), (1, 256, "pk2048")])
def test_long_runs_tensor_memory_ring(gpu_engine_factory, ch, hop, kernel):
    """The bench workload shape (stereo 48 kHz, FFT 2048, hop 512; 64 streams x 20 s = 120 k frames in one launch) and its mono
    siblings: launches this long go to the long-run instantiations (PKZ_RING of stft_pkz2048_kernel, PK_LD_RING8 / RING4 of
    stft_pk2048_kernel: contiguous columns per warp, the frame's samples in a tensor-memory ring, only the new chunk staged).
    Bit-identical to the same columns rendered in short launches (frames dealt round-robin, whole frames staged) -- i.e. runs
    that start anywhere, cross stream boundaries and include the pre-roll columns -- plus spot parity against the oracle, the
    pre-roll colour and determinism."""
    import torch
    fs, N, S, n = 48000.0, 2048, 64, 48000 * 20
    B = N // 2 + 1
    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, hop=hop, channels=ch)
    assert eng.kernel_name == kernel
    ncols = eng.columns_for(n)
    d_in = torch.empty((S, ch, n), dtype=torch.float32, device="cuda")
    eng.synth_device(d_in.data_ptr(), S, ch, n, ch * n, n, kind="mix", seed=2)
    d_pix = torch.empty((S, ncols, B), dtype=torch.int32, device="cuda")
    eng.render_device(d_in.data_ptr(), S, n, ch * n, n, 0, ncols, d_pix.data_ptr(), None)
    eng.sync()
    assert S * ncols >= 16 * 148 * 12, "the launch must be long enough for the ring instantiation"
    floor = int(_floor_pixel(eng).view(np.int32))
    assert bool((d_pix[:, 0] == floor).all()), "column 0 must be the -110 dB pre-roll colour"
    # short launches (two streams, a few hundred columns each: below the ring threshold) must reproduce it bit for bit
    d_part = torch.empty((2, 300, B), dtype=torch.int32, device="cuda")
    for s0, c0 in ((0, 0), (17, 911), (62, ncols - 300)):
        eng.render_device(d_in[s0].data_ptr(), 2, n, ch * n, n, c0, 300, d_part.data_ptr(), None)
        eng.sync()
        assert torch.equal(d_part, d_pix[s0:s0 + 2, c0:c0 + 300]), f"ring rendering differs from short launches at stream {s0}, column {c0}"
    rng = np.random.default_rng(2)
    for _ in range(3):
        s, c0 = int(rng.integers(S)), int(rng.integers(8, ncols - 8))
        _spot_check(eng, d_in[s], n, N, hop, c0, 6, d_pix[s, c0:c0 + 6], lambda db: db[:, ::-1], f"stream {s}")
    first = _checksum(d_pix)
    eng.render_device(d_in.data_ptr(), S, n, ch * n, n, 0, ncols, d_pix.data_ptr(), None)
    eng.sync()
    assert _checksum(d_pix) == first


def test_cfg4_full_size_1024_streams_60s(gpu_engine_factory):
    """1024-channel 48 kHz, FFT 2048, hop 256, 60 s per channel: 11.52 M frames, 11.8 GB in, 47.2 GB out (in 8 slabs)."""
    import torch
    fs, N, hop, S, n = 48000.0, 2048, 256, 1024, 48000 * 60
    B = N // 2 + 1
    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, hop=hop, channels=1)
    assert eng.kernel_name == "pk2048"
    ncols = eng.columns_for(n)
    assert ncols == n // hop + 1 == 11251
    G = 128  # streams per slab
    d_in = torch.empty((S, 1, n), dtype=torch.float32, device="cuda")
    eng.synth_device(d_in.data_ptr(), S, 1, n, n, n, kind="mix", seed=4)
    d_pix = torch.empty((G, ncols, B), dtype=torch.int32, device="cuda")
    d_half = torch.empty((G, ncols // 2, B), dtype=torch.int32, device="cuda")
    floor = int(_floor_pixel(eng).view(np.int32))
    sums, frames = [], 0
    rng = np.random.default_rng(4)
    for g in range(S // G):
        base = d_in[g * G]
        eng.render_device(base.data_ptr(), G, n, n, n, 0, ncols, d_pix.data_ptr(), None)
        eng.sync()
        frames += G * ncols
        sums.append(_checksum(d_pix))
        assert bool((d_pix[:, 0] == floor).all()), "column 0 must be the -110 dB pre-roll colour"
        if g in (0, 5):
            # the multi-GPU column partition: second half rendered on its own (re-reading its halo) is bit-identical
            c1 = ncols - ncols // 2
            eng.render_device(base.data_ptr(), G, n, n, n, c1, ncols // 2, d_half.data_ptr(), None)
            eng.sync()
            assert torch.equal(d_half, d_pix[:, c1:]), "column-range shard differs from the whole rendering"
            s, c0 = int(rng.integers(G)), int(rng.integers(16, ncols - 16))
            _spot_check(eng, d_in[g * G + s], n, N, hop, c0, 6, d_pix[s, c0:c0 + 6], lambda db: db[:, ::-1], f"cfg4 slab {g}")
    assert frames == 11_520_000 + 1024  # 11 250 hops + the column that ends at the last sample, per stream
    # determinism: checksum of checksums reproduces
    eng.render_device(d_in[0].data_ptr(), G, n, n, n, 0, ncols, d_pix.data_ptr(), None)
    eng.sync()
    assert _checksum(d_pix) == sums[0]
    assert len(set(sums)) == len(sums), "independent streams must not produce identical slabs"


def test_cfg3_full_size_one_hour_fft16384(gpu_engine_factory):
    """mono 96 kHz, FFT 16384, hop 4096, Blackman-Harris, 1 h: 84 375 columns of 8193 rows."""
    import torch
    fs, N, hop, n = 96000.0, 16384, 4096, 96000 * 3600
    B = N // 2 + 1
    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, hop=hop, channels=1, window="blackmanharris")
    ncols = eng.columns_for(n)
    assert ncols == n // hop + 1 == 84376
    d_in = torch.empty((1, 1, n), dtype=torch.float32, device="cuda")
    eng.synth_device(d_in.data_ptr(), 1, 1, n, n, n, kind="mix", seed=3)
    d_pix = torch.empty((ncols, B), dtype=torch.int32, device="cuda")
    eng.render_device(d_in.data_ptr(), 1, n, n, n, 0, ncols, d_pix.data_ptr(), None)
    eng.sync()
    total = _checksum(d_pix)
    # 8 column shards (the 8-GPU partition of one long stream), each re-reading its N-hop halo
    parts = torch.empty_like(d_pix)
    for g in range(8):
        c0, c1 = ncols * g // 8, ncols * (g + 1) // 8
        eng.render_device(d_in.data_ptr(), 1, n, n, n, c0, c1 - c0, parts[c0:c1].data_ptr(), None)
    eng.sync()
    assert torch.equal(parts, d_pix) and _checksum(parts) == total
    rng = np.random.default_rng(3)
    for c0 in (0, int(rng.integers(8, ncols - 8)), ncols - 4):
        _spot_check(eng, d_in[0], n, N, hop, c0, 4, d_pix[c0:c0 + 4], lambda db: db[:, ::-1], f"cfg3 col {c0}")


def test_cfg5_full_size_24h_fft65536_log_rows(gpu_engine_factory):
    """mono 192 kHz, FFT 65536, hop 1024, log-frequency max-pool to 1080 rows, 24 h of audio: 16.2 M columns.
    The input (66 GB) stays resident; the output is rendered slab by slab into a reused 1 h buffer and check-summed."""
    import torch
    fs, N, hop, R = 192000.0, 65536, 1024, 1080
    free, _ = torch.cuda.mem_get_info()
    hours = 24 if free > 90e9 else max(1, int((free - 12e9) / 2.9e9))
    n = int(fs) * 3600 * hours
    eng = gpu_engine_factory(sample_rate=fs, fft_size=N, hop=hop, channels=1, row_map="log_maxpool", rows=R, fmin=20.0, fmax=96000.0)
    ncols = eng.columns_for(n)
    assert ncols == n // hop + 1
    if hours == 24:
        assert ncols == 16_200_001
    d_in = torch.empty((1, 1, n), dtype=torch.float32, device="cuda")
    eng.synth_device(d_in.data_ptr(), 1, 1, n, n, n, kind="mix", seed=5)
    slab = 192000 * 3600 // hop  # one hour of columns
    d_pix = torch.empty((slab + 1, R), dtype=torch.int32, device="cuda")
    d_edge = torch.empty((16, R), dtype=torch.int32, device="cuda")
    blo, bhi = np.zeros(R, np.int32), np.zeros(R, np.int32)
    eng.lib.jade_log_rows(C.c_float(fs), N, R, C.c_float(20.0), C.c_float(96000.0), blo.ctypes.data, bhi.ctypes.data)

    def pooled_rows(odb):
        return np.stack([odb[:, blo[r]:bhi[r]].max(axis=1) for r in range(R)], axis=1)[:, ::-1]

    sums, done = [], 0
    prev_tail = None
    rng = np.random.default_rng(5)
    for h in range(hours):
        c0 = h * slab
        c1 = min(ncols, c0 + slab + (1 if h == hours - 1 else 0))
        eng.render_device(d_in.data_ptr(), 1, n, n, n, c0, c1 - c0, d_pix.data_ptr(), None)
        eng.sync()
        done += c1 - c0
        sums.append(_checksum(d_pix[:c1 - c0]))
        if h in (0, hours // 2, hours - 1):
            # slab boundary: the 16 columns straddling it, rendered on their own, equal the two slabs' columns
            if prev_tail is not None:
                eng.render_device(d_in.data_ptr(), 1, n, n, n, c0 - 8, 16, d_edge.data_ptr(), None)
                eng.sync()
                assert torch.equal(d_edge[8:], d_pix[:8]) and torch.equal(d_edge[:8], prev_tail)
            k0 = int(rng.integers(70, c1 - c0 - 8))
            _spot_check(eng, d_in[0], n, N, hop, c0 + k0, 3, d_pix[k0:k0 + 3], pooled_rows, f"cfg5 hour {h}")
        prev_tail = d_pix[c1 - c0 - 8:c1 - c0].clone()
    assert done == ncols
    assert len(set(sums)) == len(sums)
