"""Parity criteria of the north star, written once.

magnitude M = sqrt(P):  |M_gpu - M_ref| <= 1e-4 * M_ref + 8 * eps32 * sqrt(log2(N) * sum(P_ref) / N)
dB                   :  <= 1e-3 dB for bins within 80 dB of the frame peak
palette index        :  identical unless the oracle dB lies within 1e-3 dB of an index edge Min + i / AccessMult
"""
import numpy as np

EPS32 = float(np.finfo(np.float32).eps)


def power_from_db(db):
    return np.maximum(10.0 ** (db.astype(np.float64) / 10.0) - 1e-11, 0.0)


def check_db(db_gpu, db_ref, N, label=""):
    assert db_gpu.shape == db_ref.shape, (db_gpu.shape, db_ref.shape)
    pg, pr = power_from_db(db_gpu), power_from_db(db_ref)
    mg, mr = np.sqrt(pg), np.sqrt(pr)
    floor = 8 * EPS32 * np.sqrt(np.log2(N) * pr.sum(axis=-1, keepdims=True) / N)
    # dB values are float32: their own quantisation (1 ulp at |dB|~128 is 7.6e-6 dB -> 1.8e-6 relative in M)
    tol = 1e-4 * mr + floor + 4e-6 * mr + 1e-9
    bad = np.abs(mg - mr) > tol
    assert not bad.any(), f"{label}: {bad.sum()} magnitudes out of tolerance, worst {np.max(np.abs(mg - mr) / tol):.2f}x"
    peak = db_ref.max(axis=-1, keepdims=True)
    near = db_ref >= peak - 80.0
    # bins near the float32 FFT noise floor are exempt from the absolute dB criterion (covered by the magnitude one)
    significant = near & (mr > 50 * floor)
    d = np.abs(db_gpu - db_ref)[significant]
    if d.size:
        assert d.max() <= 1e-3, f"{label}: dB differs by {d.max():.2e} within 80 dB of the peak"


def check_pixels(pix_gpu, pix_ref, db_ref_rows, pmin, pmax, ncolors, label=""):
    """pix_* [..., rows]; db_ref_rows: oracle dB per row in the same order as the pixels."""
    assert pix_gpu.shape == pix_ref.shape
    diff = pix_gpu != pix_ref
    if not diff.any():
        return 0
    mult = np.float32(ncolors) / (np.float32(pmax) - np.float32(pmin))
    v = db_ref_rows.astype(np.float64)
    pos = (np.clip(v, pmin, pmax) - pmin) * float(mult)
    dist_db = np.abs(pos - np.round(pos)) / float(mult)
    near_edge = (dist_db <= 1e-3) | (np.abs(v - pmax) <= 1e-3) | (np.abs(v - pmin) <= 1e-3)
    offenders = diff & ~near_edge
    assert not offenders.any(), f"{label}: {offenders.sum()} palette indices differ away from a bin edge"
    return int(diff.sum())
