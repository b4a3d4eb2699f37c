"""Parity criteria of the north star, written once.

magnitude M = sqrt(P):  |M_gpu - M_ref| <= 1e-4 * M_ref + 8 * eps32 * sqrt(log2(N) * sum(P_ref) / N)
dB                   :  <= 1e-3 dB for bins within 80 dB of the frame peak
palette index        :  identical unless the oracle dB lies within the dB tolerance of an index edge Min + i / AccessMult,
                        the dB tolerance of a bin being 1e-3 dB or, for bins near the float32 noise floor of their frame,
                        what the magnitude criterion above allows there (whichever is larger)
"""
import numpy as np

EPS32 = float(np.finfo(np.float32).eps)


def q_from_db(db):
    """Q = p + 1e-11 (the quantity the reference takes the log of, Spectrogram.cpp:36,107)."""
    return 10.0 ** (db.astype(np.float64) / 10.0)


def magnitude_tolerance(db_ref, N):
    """(mr, tol, floor): oracle magnitude sqrt(Q), the allowed |M_gpu - M_ref| per bin and its noise-floor term."""
    qr = q_from_db(db_ref)
    mr = np.sqrt(qr)
    floor = 8 * EPS32 * np.sqrt(np.log2(N) * np.maximum(qr - 1e-11, 0).sum(axis=-1, keepdims=True) / N)
    # the compared dB values are float32: 1 ulp at |dB| in [64,128) is 7.6e-6 dB = 0.9e-6 relative in sqrt(Q)
    return mr, 1e-4 * mr + floor + 4e-6 * mr, floor


def check_db(db_gpu, db_ref, N, label=""):
    assert db_gpu.shape == db_ref.shape, (db_gpu.shape, db_ref.shape)
    mg = np.sqrt(q_from_db(db_gpu))
    mr, tol, floor = magnitude_tolerance(db_ref, N)
    bad = np.abs(mg - mr) > tol
    assert not bad.any(), f"{label}: {bad.sum()} magnitudes out of tolerance, worst {np.max(np.abs(mg - mr) / tol):.2f}x"
    # the same criterion in dB (1e-4 relative == 8.7e-4 dB): bins whose magnitude is far enough above the float32 FFT
    # noise floor that the floor term is negligible must agree to 1e-3 dB
    clear = mr > 2e4 * floor
    d = np.abs(db_gpu - db_ref)[clear]
    if d.size:
        assert d.max() <= 1e-3, f"{label}: dB differs by {d.max():.2e} on bins well above the noise floor"


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
    # dB tolerance per bin: 1e-3 dB, or what the magnitude criterion allows near the frame's float32 noise floor (identity
    # row maps only: rows = N/2 + 1, so that the frame sums of check_db can be formed from the rows)
    tol_db = np.full(v.shape, 1e-3)
    nb = v.shape[-1] - 1
    if nb >= 32 and (nb & (nb - 1)) == 0:
        mr, tol, _ = magnitude_tolerance(db_ref_rows, 2 * nb)
        tol_db = np.maximum(tol_db, 20.0 * np.log10(1.0 + tol / mr))
    near_edge = (dist_db <= tol_db) | (np.abs(v - pmax) <= tol_db) | (np.abs(v - pmin) <= tol_db)
    offenders = diff & ~near_edge
    assert not offenders.any(), f"{label}: {offenders.sum()} palette indices differ away from a bin edge"
    return int(diff.sum())
