"""ctypes access to the CPU SIMT emulator build of the kernels (tests/emu).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import pathlib
import subprocess

import numpy as np

from jadespectrogram_b200._capi import JadeConfig

ROOT = pathlib.Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "emu"
SO = EMU_DIR / "libjade_emu.so"
SRCS = [EMU_DIR / "emu_harness.cpp", EMU_DIR / "cuda_emu.h", ROOT / "jadespectrogram_b200/csrc/jade_kernels.cuh",
        ROOT / "jadespectrogram_b200/csrc/jade_fft_regs.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pkz.cuh", ROOT / "jadespectrogram_b200/csrc/jade_tmem.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk3.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk_cluster3.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk_cluster.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk_cta.cuh", ROOT / "jadespectrogram_b200/csrc/jade_pk_small.cuh", ROOT / "jadespectrogram_b200/csrc/jade_host_tables.cpp"]


def build():
    if SO.exists() and all(SO.stat().st_mtime >= s.stat().st_mtime for s in SRCS):
        return
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-I", str(EMU_DIR),
                    "-o", str(SO), str(EMU_DIR / "emu_harness.cpp"),
                    str(ROOT / "jadespectrogram_b200/csrc/jade_host_tables.cpp")], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(SO))
        _lib.emu_render.argtypes = [C.POINTER(JadeConfig), C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                    C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def render(cfg, palette, min_db, max_db, samples, first_col, ncols, rows, grid=2, want_db=True):
    """samples [nstreams][channels][nsamples] float32 -> (db [nstreams][ncols][B], pix [nstreams][ncols][rows])."""
    samples = np.ascontiguousarray(samples, np.float32)
    nstreams, ch, ns = samples.shape
    assert ch == cfg.channels
    B = cfg.fft_size // 2 + 1
    palette = np.ascontiguousarray(palette, np.int32)
    pix = np.zeros((nstreams, ncols, rows), np.uint32)
    db = np.zeros((nstreams, ncols, B), np.float32)
    r = lib().emu_render(C.byref(cfg), palette.ctypes.data, palette.size, min_db, max_db, samples.ctypes.data, nstreams, ns,
                         first_col, ncols, grid, pix.ctypes.data, db.ctypes.data if want_db else None)
    assert r == rows, (r, rows)
    return db, pix
