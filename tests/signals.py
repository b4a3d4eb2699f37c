"""Deterministic synthetic inputs shared by the tests and bench.py."""
import numpy as np

SEED = 20240601


def sweep(n, fs, f0=20.0, f1=None, amp=0.5, phase=0.0):
    f1 = 0.475 * fs if f1 is None else f1
    t = np.arange(n, dtype=np.float64) / fs
    dur = n / fs
    k = (f1 - f0) / dur
    return (amp * np.sin(2 * np.pi * (f0 * t + 0.5 * k * t * t + phase))).astype(np.float32)


def noise(n, seed=SEED, amp=0.5):
    rng = np.random.default_rng(seed)
    return (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32) * np.float32(2 * amp)


def mix(n, fs, seed=SEED):
    return (0.1 * noise(n, seed) + sweep(n, fs)).astype(np.float32)


def streams(nstreams, channels, n, fs, kind="mix", seed=SEED):
    out = np.empty((nstreams, channels, n), np.float32)
    for s in range(nstreams):
        for c in range(channels):
            sd = seed + 1000 * s + c
            if kind == "sweep":
                out[s, c] = sweep(n, fs, phase=0.25 * c + 0.1 * s)
            elif kind == "noise":
                out[s, c] = noise(n, sd)
            else:
                out[s, c] = 0.1 * noise(n, sd) + sweep(n, fs, phase=0.25 * c + 0.1 * s)
    return out
