"""Offline file harness: WAV -> image through the engine, against the oracle; state round trip."""
import json

import numpy as np
import pytest

import oracle_lib as O
import parity
import signals

pytestmark = pytest.mark.gpu


def test_wav_to_image_matches_oracle_and_state_round_trips(tmp_path):
    from scipy.io import wavfile

    from jadespectrogram_b200 import render_file
    fs, N, hop = 48000, 1024, 256
    x = signals.streams(1, 2, fs // 2, float(fs), kind="mix", seed=9)[0]
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype(np.int16)
    wav = tmp_path / "in.wav"
    wavfile.write(str(wav), fs, pcm)
    out, state = tmp_path / "out.ppm", tmp_path / "state.json"
    assert render_file.main([str(wav), str(out), "--fft", str(N), "--hop", str(hop), "--scheme", "viridis", "--range", "-80", "0",
                             "--state", str(state)]) == 0
    xq = np.ascontiguousarray(pcm.T.astype(np.float32) / 32768.0)  # what the tool read back
    img, st = render_file.render(xq, float(fs), fft=N, hop=hop, scheme="viridis", vmin=-80.0, vmax=0.0)
    ncols = (xq.shape[1] - N) // hop + 1
    assert img.shape == (N // 2 + 1, ncols)
    # the oracle's column j analyses x[j*hop - N, j*hop): the file harness has no pre-roll, so its column c is column c + N/hop
    odb, opix = O.render_batch(xq, fs=float(fs), fft_size=N, hop=hop, scheme="viridis", min_db=-80.0, max_db=0.0,
                               first_col=N // hop, ncols=ncols)
    parity.check_pixels(img.T, opix, odb[:, ::-1], -80.0, 0.0, 256, "render_file")
    # the PPM on disk is the same picture
    raw = out.read_bytes()
    header_end = raw.index(b"255\n") + 4
    rgb = np.frombuffer(raw[header_end:], np.uint8).reshape(img.shape[0], img.shape[1], 3)
    assert np.array_equal((rgb[..., 0].astype(np.uint32) << 16) | (rgb[..., 1].astype(np.uint32) << 8) | rgb[..., 2], img & 0xFFFFFF)
    # state round trip: re-rendering from the saved JSON alone reproduces the image bit for bit
    saved = json.load(open(state))
    assert saved["fft_size"] == N and saved["hop"] == hop and saved["palette_scheme"] == "viridis"
    img2, st2 = render_file.render(xq, float(fs), state=saved)
    assert np.array_equal(img, img2) and st2 == saved
