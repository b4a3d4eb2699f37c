#!/usr/bin/env python3
"""Library baseline for the cfg2-batch workload: the same path (frame, window, real FFT, |X|^2, channel mean, dB, palette
index, ARGB gather, flipped column store) written with torch eager ops on top of cuFFT (torch.stft), device-resident, timed
with CUDA events -- the "recompiled library kernels" comparison for the fused kernel of jade_pkz.cuh.  Not part of the product
or the tests; prints one JSON line.   usage: cufft_baseline.py [streams=256] [seconds=20] [chunk=32] [steps=5]"""
import json, sys, torch

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 32
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
fs, N, hop, C, npal = 48000, 2048, 512, 2, 256
dev = torch.device("cuda:0")
ns = int(seconds * fs)
ncols = (ns - N) // hop + 1
g = torch.Generator(device=dev).manual_seed(1)
x = 0.1 * torch.randn(streams, C, ns, device=dev, generator=g)
n = torch.arange(N, device=dev, dtype=torch.float64)
w = (0.5 - 0.5 * torch.cos(2 * torch.pi * n / N))
w = (w / torch.sqrt(torch.mean(w * w))).float()  # unit-RMS Hann (Spectrogram.cpp:239-293)
pal = (torch.arange(npal + 1, device=dev, dtype=torch.int32) * 65793) | -16777216  # any 256-entry ARGB table
vmin, vmax = -50.0, 50.0
mult = npal / (vmax - vmin)
out = torch.empty(streams, ncols, N // 2 + 1, dtype=torch.int32, device=dev)

def render():
    for s0 in range(0, streams, chunk):
        xs = x[s0:s0 + chunk].reshape(-1, ns)
        Z = torch.stft(xs, n_fft=N, hop_length=hop, window=w, center=False, return_complex=True)  # [chunk*C, bins, cols] (cuFFT)
        p = (Z.real * Z.real + Z.imag * Z.imag).reshape(-1, C, N // 2 + 1, ncols).mean(1)
        db = 10.0 * torch.log10(p + 1e-11)
        idx = ((db - vmin) * mult).clamp_(0, npal).to(torch.int64)
        out[s0:s0 + chunk] = pal[idx].flip(1).transpose(1, 2)  # [stream, col, row], row 0 = highest bin

for _ in range(2):
    render()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    render()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
frames = streams * ncols

def fft_only():  # cuFFT alone (framing + window + R2C, complex spectrum written to HBM): what the library part costs
    for s0 in range(0, streams, chunk):
        torch.stft(x[s0:s0 + chunk].reshape(-1, ns), n_fft=N, hop_length=hop, window=w, center=False, return_complex=True)

fft_only()
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    fft_only()
e1.record()
torch.cuda.synchronize()
ms_fft = e0.elapsed_time(e1) / steps
print(json.dumps({"impl": "torch eager + cuFFT (torch.stft)", "workload": "cfg2-batch", "frames_per_pass": frames, "ms_per_pass": ms,
                  "frames_per_s": frames / ms * 1e3, "algorithmic_GBps": frames * 8196 / ms / 1e6, "chunk_streams": chunk,
                  "stft_only_ms_per_pass": ms_fft, "stft_only_frames_per_s": frames / ms_fft * 1e3}))
