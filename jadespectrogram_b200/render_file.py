#!/usr/bin/env python3
"""Offline file harness: WAV in -> spectrogram image out through the B200 engine.

    python -m jadespectrogram_b200.render_file in.wav out.png --fft 2048 --hop 512 --window hann --scheme jade \\
           --range -50 50 [--rows 1080 --row-map log_maxpool --fmin 20 --fmax 20000] [--gpus 2] [--state cfg.json]

The picture is what the plugin's SpectrogramComponent would have drawn for the same audio (Spectrogram.cpp:590-731 in
scroll mode with a ring as long as the file): x = column (time), y = H-1-bin (low frequencies at the bottom), colours from
CColorPalette.  All arithmetic runs in libjade_gpu.so (jade_render_batch / jade_render_batch_multi); this module only
reads the file, hands buffers over and writes the image.  `--state` saves / restores the engine configuration as JSON --
the analogue of the plugin's getStateInformation / setStateInformation round trip (PluginProcessor.cpp:164-185).
"""
import argparse
import ctypes as C
import json
import sys

import numpy as np

from . import _capi
from .engine import Engine, JadeError, default_config

STATE_FIELDS = [f for f, _ in _capi.JadeConfig._fields_]


def read_wav(path):
    """-> (fs, planar float32 [channels][n]).  PCM 8/16/24/32 and float WAV via scipy.io.wavfile."""
    from scipy.io import wavfile
    fs, data = wavfile.read(path)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    elif np.issubdtype(data.dtype, np.integer):
        x = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    else:
        x = data.astype(np.float32)
    return float(fs), np.ascontiguousarray(x.T)


def config_to_state(cfg, scheme, ncolors, invert, vmin, vmax):
    d = {f: getattr(cfg, f) for f in STATE_FIELDS}
    d.update(palette_scheme=scheme, palette_colors=ncolors, palette_invert=bool(invert), min_db=vmin, max_db=vmax)
    return d


def state_to_config(d):
    cfg = _capi.JadeConfig()
    for f in STATE_FIELDS:
        if f in d:
            setattr(cfg, f, d[f])
    return cfg


def render(samples, fs, *, fft=2048, hop=512, window="hann", mix="absmean", scheme="jade", ncolors=256, invert=False,
           vmin=-50.0, vmax=50.0, row_map="identity", rows=0, fmin=0.0, fmax=20000.0, gpus=1, state=None):
    """samples planar [channels][n] -> (image uint32 [rows][columns] ARGB, state dict)."""
    ch, n = samples.shape
    if state is not None:
        cfg = state_to_config(state)
        scheme, ncolors, invert = state["palette_scheme"], state["palette_colors"], state["palette_invert"]
        vmin, vmax = state["min_db"], state["max_db"]
    else:
        cfg = default_config(sample_rate=fs, fft_size=fft, hop=hop, channels=ch, window=window, mix_mode=mix, row_map=row_map,
                             rows=rows, fmin=fmin, fmax=fmax, preroll=0, ring_columns=8)
    engines = [Engine(g, cfg) for g in range(gpus)]
    try:
        for e in engines:
            e.set_palette_scheme(scheme, ncolors, invert)
            e.set_value_range(vmin, vmax)
        e0 = engines[0]
        ncols = e0.columns_for(n)
        if ncols <= 0:
            raise JadeError(f"{n} samples are shorter than one analysis frame of {e0.N}")
        pix = np.empty((1, ncols, e0.R), np.uint32)
        x = np.ascontiguousarray(samples[None], np.float32)
        if gpus == 1:
            e0.render_batch(x, out_pix=pix)
        else:  # one stream: column ranges are sharded over the GPUs, each re-reads its N-hop halo (no exchange)
            arr = (C.c_void_p * gpus)(*[e.h for e in engines])
            rc = e0.lib.jade_render_batch_multi(arr, gpus, x.ctypes.data, 1, n, 0, ncols, pix.ctypes.data, None)
            if rc != 0:
                raise JadeError(f"jade_render_batch_multi failed ({rc}): {e0.lib.jade_last_error(None).decode()}")
        st = config_to_state(e0.get_config(), scheme, ncolors, invert, vmin, vmax)
        return np.ascontiguousarray(pix[0].T), st  # [rows][columns]: x = time, row 0 = highest frequency
    finally:
        for e in engines:
            e.close()


def write_image(path, argb):
    """argb uint32 [rows][cols] (0xAARRGGBB).  .png via Pillow, .ppm / .raw written directly."""
    h, w = argb.shape
    rgb = np.stack([(argb >> 16) & 255, (argb >> 8) & 255, argb & 255], axis=-1).astype(np.uint8)
    if path.endswith(".raw"):
        np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], axis=-1).tofile(path)  # RGBA8
    elif path.endswith(".ppm"):
        with open(path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h))
            f.write(rgb.tobytes())
    else:
        from PIL import Image
        Image.fromarray(rgb, "RGB").save(path)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("wav")
    ap.add_argument("out", help=".png / .ppm / .raw (RGBA8)")
    ap.add_argument("--fft", type=int, default=2048)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--window", default="hann", choices=list(_capi.WIN))
    ap.add_argument("--mix", default="absmean", choices=list(_capi.MIX))
    ap.add_argument("--scheme", default="jade", choices=list(_capi.PAL))
    ap.add_argument("--colors", type=int, default=256)
    ap.add_argument("--invert", action="store_true")
    ap.add_argument("--range", type=float, nargs=2, default=[-50.0, 50.0], metavar=("MIN_DB", "MAX_DB"))
    ap.add_argument("--row-map", default="identity", choices=list(_capi.ROWS))
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--fmin", type=float, default=0.0)
    ap.add_argument("--fmax", type=float, default=20000.0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--state", help="JSON file: restored if it exists, written after rendering")
    a = ap.parse_args(argv)
    fs, x = read_wav(a.wav)
    state = None
    if a.state:
        try:
            state = json.load(open(a.state))
        except FileNotFoundError:
            state = None
    img, st = render(x, fs, fft=a.fft, hop=a.hop, window=a.window, mix=a.mix, scheme=a.scheme, ncolors=a.colors,
                     invert=a.invert, vmin=a.range[0], vmax=a.range[1], row_map=a.row_map, rows=a.rows, fmin=a.fmin, fmax=a.fmax,
                     gpus=a.gpus, state=state)
    write_image(a.out, img)
    if a.state:
        json.dump(st, open(a.state, "w"), indent=1)
    print(f"{a.wav}: {x.shape[0]} ch, {x.shape[1]} samples @ {fs:.0f} Hz -> {a.out} {img.shape[1]} x {img.shape[0]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
